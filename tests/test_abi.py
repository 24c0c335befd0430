"""The C-ABI library loads and exports exactly what include/phylo_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from phylo_utils_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "phylo_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(phb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_functions():
    names = declared_functions()
    assert "phb_create" in names and "phb_compute_partials" in names and len(names) >= 25


def test_every_declared_symbol_is_exported():
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(handle, name), "libphylo_b200.so does not export " + name


def test_python_binding_table_matches_header():
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_header_cites_reference_interfaces():
    text = open(HEADER).read()
    for cite in ("tree_model.py:160-176", "numba_likelihood_engine.py:10-46", "abstract.py:49-59",
                 "c_discrete_gamma.c:285-321", "tree_model.py:178-217"):
        assert cite in text


def test_version_and_status_names():
    lib = _lib.lib()
    assert lib.phb_version() == 100
    assert lib.phb_status_name(0) == b"PHB_OK"
    assert lib.phb_status_name(3) == b"PHB_ERR_NO_DEVICE"


def test_workspace_size_of_headline_config():
    # 1000 taxa x 1M patterns x 4 categories x 4 states: 998 internal nodes x 128 MB + scalers + tips
    lib = _lib.lib()
    n = lib.phb_workspace_bytes(1000, 1000000, 4, 4, 0)
    partials = 998 * 1000000 * 16 * 8
    assert partials < n < partials * 1.06
    assert n < 180e9            # fits one B200
    assert lib.phb_workspace_bytes(1000, 1000000, 4, 4, _lib.PHB_FLAG_NO_PARTIALS) < 1.4e9
    assert lib.phb_workspace_bytes(1, 10, 4, 4, 0) == 0          # invalid shape
    assert lib.phb_workspace_bytes(10, 10, 4, 65, 0) == 0


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_a_device():
    from phylo_utils_b200.engine import LikelihoodEngine
    with pytest.raises(RuntimeError) as exc:
        LikelihoodEngine(4, 10, 4, 4)
    assert "no CPU fallback" in str(exc.value) or "NO_DEVICE" in str(exc.value)
    import numpy as np
    from phylo_utils_b200.likelihood import clv
    p = np.eye(4)[None]
    with pytest.raises(RuntimeError):
        clv(p, p, np.ones((3, 1, 4)), np.ones((3, 1, 4)), np.zeros((3, 1)), np.zeros((3, 1)), np.zeros((3, 1)))


def test_bad_arguments_are_value_errors():
    from phylo_utils_b200.engine import LikelihoodEngine
    with pytest.raises(ValueError):
        LikelihoodEngine(1, 10, 4, 4)
    with pytest.raises(ValueError):
        LikelihoodEngine(4, 10, 4, 100)


def test_pack_codes_is_host_only_and_exact():
    """phb_pack_codes needs no GPU: even pattern in the low nibble, odd length padded with a zero nibble."""
    import numpy as np
    from phylo_utils_b200 import LikelihoodEngine
    rng = np.random.default_rng(0)
    for ntax, nsite in [(1, 1), (3, 2), (4, 7), (5, 64), (2, 1001)]:
        codes = rng.integers(0, 16, size=(ntax, nsite)).astype(np.uint8)
        packed = LikelihoodEngine.pack_codes(codes)
        assert packed.shape == (ntax, (nsite + 1) // 2) and packed.dtype == np.uint8
        padded = np.concatenate([codes, np.zeros((ntax, nsite % 2), dtype=np.uint8)], axis=1)
        assert np.array_equal(packed, padded[:, 0::2] | (padded[:, 1::2] << 4))
        assert np.array_equal(packed & 15, padded[:, 0::2]) and np.array_equal(packed >> 4, padded[:, 1::2])
    with pytest.raises(ValueError):
        LikelihoodEngine.pack_codes(np.full((2, 4), 16, dtype=np.uint8))      # does not fit in four bits
    with pytest.raises(ValueError):
        LikelihoodEngine.pack_codes(np.zeros(5, dtype=np.uint8))              # not (n_tips, n_patterns)


def test_compression_and_context_entry_points_fail_loudly_without_a_gpu():
    """No CPU fallback anywhere: without a device the calls return PHB_ERR_NO_DEVICE (RuntimeError), they do not compute."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from phylo_utils_b200.alignment.alignment import compress_codes_gpu
    with pytest.raises(RuntimeError):
        compress_codes_gpu(np.zeros((3, 10), dtype=np.uint8))
    from phylo_utils_b200 import LikelihoodEngine
    with pytest.raises(RuntimeError):
        LikelihoodEngine(4, 10, 4, 4)


def test_code_packing_helpers_round_trip():
    import numpy as np
    from phylo_utils_b200 import LikelihoodEngine
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 64, 1001):
        codes = rng.integers(0, 8, size=(5, n)).astype(np.uint8)
        low, high = LikelihoodEngine.split_codes(codes)
        assert low.shape == (5, (n + 3) // 4) and high.shape == (5, (n + 7) // 8)
        s = np.arange(n)
        got = ((low[:, s // 4] >> (2 * (s % 4))) & 3) | (((high[:, s // 8] >> (s % 8)) & 1) << 2)
        assert np.array_equal(got, codes)
        packed = LikelihoodEngine.pack_codes(codes)
        assert np.array_equal((packed[:, s // 2] >> (4 * (s % 2))) & 15, codes)
    with pytest.raises(ValueError):
        LikelihoodEngine.split_codes(np.full((2, 3), 8, dtype=np.uint8))
    with pytest.raises(ValueError):
        LikelihoodEngine.pack_codes(np.full((2, 3), 16, dtype=np.uint8))


def test_build_script_lists_every_source_and_header():
    """A translation unit or header that exists under csrc/ but is missing from build.py would silently drop kernels from the
    library (the lnL-only walk is instantiated in four units) or leave stale objects after a header edit."""
    import importlib.util
    csrc = os.path.join(ROOT, "phylo_utils_b200", "csrc")
    spec = importlib.util.spec_from_file_location("phb_build", os.path.join(csrc, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    on_disk = sorted(f for f in os.listdir(csrc) if f.endswith((".cu", ".cpp")))
    assert on_disk == sorted(mod.CUDA_SOURCES + mod.HOST_SOURCES)
    headers = sorted(f for f in os.listdir(csrc) if f.endswith(".cuh"))
    assert headers == sorted(h for h in mod.HEADERS if not os.path.isabs(h))
    assert "-lineinfo" in mod.NVCC_FLAGS and "arch=compute_100a,code=sm_100a" in mod.ARCH
