"""The built library must contain the Blackwell instruction sequences the design claims (no GPU needed: cuobjdump
disassembles the sm_100a cubin).  UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM =
tcgen05.ld / st (tensor memory), UTCBAR = tcgen05.commit, HMMA = legacy mma.sync (allowed in the <= 16-row GEMM only)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fireredtts2_b200", "libfrt2_b200.so")


@pytest.fixture(scope="module")
def sass():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    txt = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)
    return {p.split("\n", 1)[0].strip(): p for p in parts[1:]}


def _fn(sass, substr):
    hits = [b for n, b in sass.items() if substr in n]
    assert hits, f"kernel {substr} not found in the library"
    return hits[0]


def _count(body, mnemonic):
    return len(re.findall(r"(?<![\w.])" + re.escape(mnemonic) + r"(?![\w])", body))


def test_every_kernel_is_compiled_for_sm_100a_only(sass):
    """One code path: every translation unit's cubin is sm_100a and there is no PTX for a JIT to retarget (the only other
    ELF is nvcc's empty device-link stub of the final `nvcc -shared`, which holds no kernel)."""
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([exe, "-lelf", LIB], capture_output=True, text=True).stdout
    files = re.findall(r"ELF file\s+\d+:\s+(\S+)", out)
    units = [f for f in files if not f.startswith("libfrt2_b200.")]
    assert len(units) >= 7 and all(f.endswith(".sm_100a.cubin") for f in units), files
    ptx = subprocess.run([exe, "-lptx", LIB], capture_output=True, text=True)
    assert "PTX file" not in ptx.stdout
    assert all("EF_CUDA_SM100" in b[:600] for b in sass.values())


def test_no_shelved_kernels_ship(sass):
    """Kernels that are off the product path do not ship: one tcgen05 attention kernel, one skinny GEMM family, no
    persistent step interpreter."""
    names = " ".join(sass)
    for gone in ("attention_tc_kernel", "attention_tcp_kernel", "gemm_skinny_fma_kernel", "stream_step_kernel"):
        assert gone not in names, gone


def test_gemm_kernels_run_on_tcgen05_with_tma(sass):
    for name in ("gemm_tc2_kernel", "gemm_tc_kernelILi256", "gemm_tc_kernelILi128"):
        b = _fn(sass, name)
        assert _count(b, "UTCHMMA") >= 4 and _count(b, "UTCBAR") >= 1, name      # tcgen05.mma + commit
        assert _count(b, "UTMALDG") >= 2 and _count(b, "UTMASTG") >= 1, name      # TMA loads of A and W, TMA store of C
        assert _count(b, "LDTM") >= 1, name                                        # accumulators read back from tensor memory
        assert _count(b, "HMMA") == 0, name                                        # not a recompiled mma.sync kernel
    assert ".2CTA" in _fn(sass, "gemm_tc2_kernel")                                 # cta_group::2


def test_attention_keeps_p_in_tensor_memory(sass):
    # product kernels: persistent, head dim 64 (3 query tiles, a quarter of the exponentials on the FMA pipe) and 128
    for name, min_ex2 in (("attention_t4_kernelILi64ELi3ELb1ELb0ELi2", 48), ("attention_t4_kernelILi128ELi2ELb1ELb0ELi2", 48),
                          ("attention_t3_kernel", 64)):   # t3: the one predecessor kept for A/B (FRT2_ATTN_VER=3)
        b = _fn(sass, name)
        assert _count(b, "UTCHMMA") >= 8 and _count(b, "UTMALDG") >= 3, name
        assert _count(b, "STTM") >= 1 and _count(b, "LDTM") >= 2, name             # P written to / S, O read from TMEM
        assert _count(b, "MUFU.EX2") >= min_ex2 and _count(b, "HMMA") == 0, name
    b = _fn(sass, "attention_t4_kernelILi64ELi3ELb1ELb0ELi2")
    assert _count(b, "FFMA2") >= 32 + 8 * 4                                        # scale/subtract + the polynomial 2^x pairs


def test_token_step_kernels(sass):
    b = _fn(sass, "gemm_skinny_kernelILi8ELi1")
    assert _count(b, "HMMA") >= 2 and _count(b, "LDGSTS") >= 1                     # mma.sync + cp.async staging
    assert _count(_fn(sass, "state_roll_kernel"), "STG.E.128") >= 1 and _count(_fn(sass, "state_reset_kernel"), "STG.E.128") >= 1
    assert _count(_fn(sass, "overlap_add_vec4_kernel"), "STG.E.128") >= 1
    assert _count(_fn(sass, "rvq_encode_kernel"), "FFMA") >= 100


def test_encode_side_kernels(sass):
    """encoder.cu: the bandwidth kernels move 16-byte vectors, the log-mel kernel is an fp32 FMA loop; the encoders'
    contractions and attention are the decode path's tcgen05 kernels (no kernel of their own)."""
    assert _count(_fn(sass, "silu_mul_kernel"), "LDG.E.EF.128") >= 2 and _count(_fn(sass, "silu_mul_kernel"), "STG.E.128") >= 1
    assert _count(_fn(sass, "cvt_rows_kernel"), "LDG.E.EF.128") >= 1      # streaming (evict-first) 16-byte loads
    assert _count(_fn(sass, "add_pos_kernel"), "STG.E.128") >= 1
    b = _fn(sass, "mel_power_kernel")
    assert _count(b, "FFMA") >= 8 and _count(b, "HMMA") == 0
    names = " ".join(sass)
    assert "rvq_encode_kernelILi8" in names and "rvq_encode_kernelILi16" in names and "rvq_encode_kernelILi32" in names


def test_frame_tail_kernels(sass):
    """gemm_stream.cu / frame_decoder.cu: the weight stream is 16-byte no-allocate loads feeding mma.sync (<= 8 rows: a
    tcgen05 tile would be 94 % padding and the kernel is HBM-bound), one block barrier, no local-memory spills; the
    sampler's selection rounds and the attention's reductions are warp shuffles."""
    b = _fn(sass, "gemm_stream_kernel")
    assert _count(b, "HMMA.16816.F32") >= 2 and _count(b, "LDG.E.128.CONSTANT") + _count(b, "LDG.E.128.NA.CONSTANT") + \
        len(re.findall(r"LDG\.E\.[\w.]*128", b)) >= 8
    assert _count(b, "UTCHMMA") == 0 and "STL" not in b and "LDL" not in b
    for name in ("fd_sample_kernel", "fd_attn_kernel"):
        k = _fn(sass, name)
        assert _count(k, "SHFL.BFLY") >= 5, name
    assert _count(_fn(sass, "fd_sample_kernel"), "MUFU.EX2") >= 1
