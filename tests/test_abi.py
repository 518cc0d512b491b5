"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/frt2.h declares; the host binding fails loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "frt2.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frt2_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from fireredtts2_b200 import build
    build.build()
    from fireredtts2_b200 import _native as N
    return N.load()


def test_every_declared_symbol_is_exported(lib):
    from fireredtts2_b200 import _native as N
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/frt2.h but not exported"
    assert sorted(N.SIGNATURES) == names, "ctypes SIGNATURES and include/frt2.h disagree"


def test_version_and_error_string(lib):
    assert b"sm_100a" in lib.frt2_version()
    assert isinstance(lib.frt2_last_error(), bytes)


def test_bad_arguments_are_rejected_without_touching_the_gpu(lib):
    from fireredtts2_b200 import _native as N
    h = C.c_void_p()
    assert lib.frt2_create(None, 0, C.byref(h)) == N.ERR_BAD_ARG
    cfg = N.Frt2Config(512, 1024, 16, 2048, 256, 1000, 12, 16, 240, 4)   # output_dim != embed_dim
    assert lib.frt2_create(C.byref(cfg), 0, C.byref(h)) == N.ERR_BAD_ARG
    cfg = N.Frt2Config(512, 1024, 16, 2048, 256, 1024, 12, 16, 240, 2)   # stride != 4
    assert lib.frt2_create(C.byref(cfg), 0, C.byref(h)) == N.ERR_BAD_ARG
    assert lib.frt2_finalize(None) == N.ERR_BAD_ARG
    assert lib.frt2_decode(None, None, 8, 0, 0, 0, 1, 1, 1, None, None, 0, None) == N.ERR_BAD_ARG


def test_frame_decoder_arguments_are_validated_without_a_gpu(lib):
    from fireredtts2_b200 import _native as N
    from fireredtts2_b200.frame_decoder import Frt2FdConfig
    h = C.c_void_p()
    assert lib.frt2_fd_create(None, 0, C.byref(h)) == N.ERR_BAD_ARG
    bad = Frt2FdConfig(1536, 1536, 4, 12, 5, 8960, 2048, 16, 1e6, 1e-6, 0)        # kv heads do not divide heads
    assert lib.frt2_fd_create(C.byref(bad), 0, C.byref(h)) == N.ERR_BAD_ARG
    bad = Frt2FdConfig(1536, 1530, 4, 12, 2, 8960, 2048, 16, 1e6, 1e-6, 0)        # dim not a multiple of 8
    assert lib.frt2_fd_create(C.byref(bad), 0, C.byref(h)) == N.ERR_BAD_ARG
    bad = Frt2FdConfig(1536, 1536, 4, 12, 2, 8960, 2048, 1, 1e6, 1e-6, 0)         # fewer than two codebooks
    assert lib.frt2_fd_create(C.byref(bad), 0, C.byref(h)) == N.ERR_BAD_ARG
    assert lib.frt2_fd_finalize(None) == N.ERR_BAD_ARG
    assert lib.frt2_fd_generate(None, None, 1, None, None, 0, 10, 1.0, None, None, None, None, None) == N.ERR_BAD_ARG
    assert lib.frt2_op_sample_topk(None, 1, 8, 2, 1.0, None, 0, None, None) == N.ERR_BAD_ARG


def test_frame_decoder_host_side():
    """config presets mirror modules.py, the weight-byte count of the roofline, the synthetic state dict's key set."""
    from fireredtts2_b200.frame_decoder import (FD_200M, FD_500M, FD_TINY, FrameDecoderConfig, frame_decoder_keys,
                                                synthetic_frame_decoder_state_dict)
    assert (FD_200M.dim, FD_200M.num_layers, FD_200M.num_heads, FD_200M.num_kv_heads, FD_200M.intermediate_dim) == \
        (1536, 4, 12, 2, 8960)                                                    # modules.py:5-18
    assert (FD_500M.dim, FD_500M.num_layers, FD_500M.num_heads, FD_500M.intermediate_dim) == (896, 24, 14, 4864)   # :21-34
    assert FD_200M.head_dim == 128 and FD_200M.qkv_dim == 2048
    per_layer = 1536 * 2048 + 1536 * 1536 + 3 * 1536 * 8960
    assert FD_200M.weight_bytes_per_frame() == 2 * (15 * (1536 * 1536 + 4 * per_layer) + 15 * 1536 * 2048 + 2048 * 1536)
    sd = synthetic_frame_decoder_state_dict(FD_TINY, 1)
    assert sorted(sd) == sorted(frame_decoder_keys(FD_TINY))
    assert sd["audio_head"].shape == (5, 64, 64) and sd["audio_embeddings.weight"].shape == (6 * 64, 96)
    with pytest.raises(ValueError):
        FrameDecoderConfig(dim=100, num_heads=3)


def test_frame_decoder_config_from_a_reference_shaped_model():
    """config_from_reference reads the widths off a module tree with the reference's names (Model.projection /
    audio_embeddings / codebook0_head / audio_head / decoder.layers[i].attn ... with torchtune parameter names)."""
    import torch
    from fireredtts2_b200.frame_decoder import FD_TINY, FrameDecoderB200, frame_decoder_keys, synthetic_frame_decoder_state_dict

    cfg = FD_TINY
    sd = synthetic_frame_decoder_state_dict(cfg, 2)

    class Holder(torch.nn.Module):
        pass

    def grow(root, dotted, tensor):
        *path, leaf = dotted.split(".")
        m = root
        for p in path:
            if not hasattr(m, p):
                setattr(m, p, Holder())
            m = getattr(m, p)
        setattr(m, leaf, torch.nn.Parameter(torch.from_numpy(tensor)))

    model = Holder()
    layers = torch.nn.ModuleList([Holder() for _ in range(cfg.num_layers)])
    model.decoder = Holder()
    model.decoder.layers = layers
    for k, v in sd.items():
        if k.startswith("decoder.layers."):
            _, _, i, rest = k.split(".", 3)
            grow(layers[int(i)], rest, v)
        else:
            grow(model, k, v)
    for l in layers:
        l.attn.num_heads = cfg.num_heads
    model.config = type("Cfg", (), {"audio_vocab_size": cfg.audio_vocab_size, "audio_num_codebooks": cfg.audio_num_codebooks})()
    got, got_sd = FrameDecoderB200.config_from_reference(model)
    assert got == cfg
    assert sorted(got_sd) == sorted(frame_decoder_keys(cfg))


def test_frame_decoder_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fireredtts2_b200.frame_decoder import FD_TINY, FrameDecoderB200, synthetic_frame_decoder_state_dict
    with pytest.raises(RuntimeError):
        FrameDecoderB200(FD_TINY, synthetic_frame_decoder_state_dict(FD_TINY, 0))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fireredtts2_b200.codec import RedCodecB200
    from fireredtts2_b200.config import MICRO
    from fireredtts2_b200.weights import synthetic_state_dict
    with pytest.raises(RuntimeError):
        RedCodecB200(MICRO, synthetic_state_dict(MICRO, 0))


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "fireredtts2_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_header_is_plain_c(tmp_path):
    """include/frt2.h is the drop-in boundary for non-Python hosts: it must compile as C (no C++-isms, no CUDA types) and
    a minimal C host that references every entry point must link against the library."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "frt2.h")).read()
    names = sorted(set(re.findall(r"\b(frt2_[a-z0-9_]+)\s*\(", hdr)))
    src = tmp_path / "host.c"
    body = "\n".join(f"  p[{i}] = (fn){n};" for i, n in enumerate(names))
    src.write_text('#include "frt2.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\nint main(void) {\n  fn p[%d];\n%s\n'
                   '  printf("%%s %%d\\n", frt2_version(), p[0] != 0);\n  return 0;\n}\n' % (len(names), body))
    lib_dir = os.path.join(root, "fireredtts2_b200")
    exe = tmp_path / "host"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(root, "include"), str(src),
                        "-o", str(exe), "-L", lib_dir, "-l:libfrt2_b200.so", "-Wl,-rpath," + lib_dir,
                        "-Wl,--unresolved-symbols=ignore-in-shared-libs"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_entry_point_is_documented_for_integrators():
    """INTEGRATION.md names every function include/frt2.h declares (the reference-side binding document stays complete)."""
    hdr = open(os.path.join(ROOT, "include", "frt2.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    names = sorted(set(re.findall(r"\b(frt2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 50
    assert [n for n in names if n not in doc] == []
