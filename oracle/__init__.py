"""CPU oracle of the hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy / torch-CPU restatements of the reference's algorithms (each function cites the reference file:line it follows), pinned
to golden vectors that the ``make_golden*.py`` scripts here generated from the UNMODIFIED reference (``tests/golden/``).
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
package, and only as the checker (or the CPU arm being timed) — ``fireredtts2_b200/`` never does
(``tests/test_abi.py::test_product_package_does_not_import_the_oracle``) and has no CPU fallback."""
