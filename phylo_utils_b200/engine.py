"""
Thin object wrapper over the C ABI (include/phylo_b200.h): one ``LikelihoodEngine`` = one
``phb_ctx`` = one GPU.  All arithmetic happens in libphylo_b200.so; this class only marshals
numpy arrays and owns the device workspace (a torch uint8 tensor when torch is importable, so
that torch's caching allocator, streams and ``torch.cuda.Event`` timing see the same memory and
stream; otherwise the library allocates for itself).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import lib, check, dptr, iptr

__all__ = ["LikelihoodEngine", "fp64_peak"]


def fp64_peak(device=0, tensor=False):
    """Measured fp64 throughput of ``device`` in TFLOP/s (vector DFMA pipe, or the DMMA tensor pipe)."""
    out = ctypes.c_double(0.0)
    check(lib().phb_op_fp64_peak(int(device), 1 if tensor else 0, ctypes.byref(out)))
    return out.value


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.double)


class LikelihoodEngine(object):
    def __init__(self, n_tips, n_patterns, n_cat, n_states, device=0, up_partials=False, store_partials=True,
                 use_torch=True):
        self.n_tips, self.n_patterns, self.n_cat, self.n_states = int(n_tips), int(n_patterns), int(n_cat), int(n_states)
        self.device = int(device)
        self._ctx = ctypes.c_void_p()
        self._keep = {}            # host/device buffers that must outlive the ctx
        self._lib = lib()
        flags = 0
        if up_partials:
            flags |= _lib.PHB_FLAG_UP_PARTIALS
        if not store_partials:
            flags |= _lib.PHB_FLAG_NO_PARTIALS
        self.flags = flags
        nbytes = self._lib.phb_workspace_bytes(self.n_tips, self.n_patterns, self.n_cat, self.n_states, flags)
        if nbytes == 0:
            raise ValueError("unsupported problem shape: tips={} patterns={} categories={} states={}".format(
                n_tips, n_patterns, n_cat, n_states))
        self.workspace_bytes = int(nbytes)
        ws_ptr, stream = None, None
        if use_torch:
            try:
                import torch
                if torch.cuda.is_available():
                    ws = torch.empty(self.workspace_bytes + 256, dtype=torch.uint8, device="cuda:{}".format(self.device))
                    base = ws.data_ptr()
                    ws_ptr = (base + 255) // 256 * 256
                    self._keep["workspace"] = ws
                    stream = torch.cuda.current_stream(self.device).cuda_stream
            except ImportError:
                pass
        check(self._lib.phb_create(self.device, self.n_tips, self.n_patterns, self.n_cat, self.n_states, flags,
                                   ctypes.c_void_p(ws_ptr), self.workspace_bytes if ws_ptr else 0,
                                   ctypes.c_void_p(stream), ctypes.byref(self._ctx)))

    # ---- life cycle -------------------------------------------------------------------------
    def close(self):
        if self._ctx:
            self._lib.phb_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()
        self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ok(self, status):
        check(status, self._ctx)

    def sync(self):
        self._ok(self._lib.phb_sync(self._ctx))

    @property
    def launch_count(self):
        return int(self._lib.phb_launch_count(self._ctx))

    # ---- inputs -----------------------------------------------------------------------------
    def set_tips(self, codes, lut, tip_nodes):
        """codes: uint8 (n_tips, n_patterns) numpy array (host) or a torch CUDA uint8 tensor (kept on device)."""
        lut = _f64(lut)
        tip_nodes = np.ascontiguousarray(tip_nodes, dtype=np.int32)
        if lut.ndim != 2 or lut.shape[1] != self.n_states:
            raise ValueError("lut must be (n_codes, n_states)")
        if tip_nodes.shape != (self.n_tips,):
            raise ValueError("tip_nodes must have one entry per tip")
        on_device = hasattr(codes, "data_ptr")
        if on_device:
            if tuple(codes.shape) != (self.n_tips, self.n_patterns) or not codes.is_contiguous():
                raise ValueError("device codes must be a contiguous (n_tips, n_patterns) uint8 tensor")
            self._keep["codes"] = codes
            ptr = ctypes.c_void_p(codes.data_ptr())
        else:
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            if codes.shape != (self.n_tips, self.n_patterns):
                raise ValueError("codes must be (n_tips, n_patterns)")
            ptr = ctypes.c_void_p(codes.ctypes.data)
        self._ok(self._lib.phb_set_tips(self._ctx, ptr, 1 if on_device else 0, lut.shape[0], dptr(lut), iptr(tip_nodes)))

    def set_pattern_weights(self, weights):
        if weights is None:
            self._ok(self._lib.phb_set_pattern_weights(self._ctx, None))
            return
        w = np.ascontiguousarray(weights, dtype=np.int64)
        if w.shape != (self.n_patterns,):
            raise ValueError("weights must have one entry per pattern")
        self._ok(self._lib.phb_set_pattern_weights(self._ctx, w.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))

    def set_model(self, evecs, evals, ivecs, freqs, rates, cat_weights):
        A, K = self.n_states, self.n_cat
        evecs, evals, ivecs = _f64(evecs), _f64(evals), _f64(ivecs)
        freqs, rates, cat_weights = _f64(freqs), _f64(rates), _f64(cat_weights)
        if evecs.shape != (A, A) or ivecs.shape != (A, A) or evals.shape != (A,) or freqs.shape != (A,):
            raise ValueError("eigen-system / frequencies do not match the number of states")
        if rates.shape != (K,) or cat_weights.shape != (K,):
            raise ValueError("rates / weights do not match the number of categories")
        self._ok(self._lib.phb_set_model(self._ctx, dptr(evecs), dptr(evals), dptr(ivecs), dptr(freqs), dptr(rates),
                                         dptr(cat_weights)))

    def set_mixture(self, freqs, rates, cat_weights):
        freqs, rates, cat_weights = _f64(freqs), _f64(rates), _f64(cat_weights)
        if freqs.shape != (self.n_states,) or rates.shape != (self.n_cat,) or cat_weights.shape != (self.n_cat,):
            raise ValueError("frequencies / rates / weights have the wrong length")
        self._ok(self._lib.phb_set_mixture(self._ctx, dptr(freqs), dptr(rates), dptr(cat_weights)))

    def set_schedule(self, rows, level_offsets=None):
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1, 3)
        if level_offsets is None:
            self._ok(self._lib.phb_set_schedule(self._ctx, rows.shape[0], iptr(rows), 0, None))
        else:
            lo = np.ascontiguousarray(level_offsets, dtype=np.int32)
            self._ok(self._lib.phb_set_schedule(self._ctx, rows.shape[0], iptr(rows), lo.shape[0] - 1, iptr(lo)))
        self.n_rows = rows.shape[0]

    def set_edge_lengths(self, lengths):
        lengths = _f64(lengths).reshape(-1, 2)
        if lengths.shape[0] != getattr(self, "n_rows", -1):
            raise ValueError("need one (len1, len2) pair per schedule row")
        self._ok(self._lib.phb_set_edge_lengths(self._ctx, dptr(lengths)))

    # ---- transition matrices ----------------------------------------------------------------
    def build_pmatrices(self):
        self._ok(self._lib.phb_build_pmatrices(self._ctx))

    def set_pmatrices(self, pmats):
        pmats = _f64(pmats)
        want = (getattr(self, "n_rows", 0), 2, self.n_cat, self.n_states, self.n_states)
        if pmats.shape != want:
            raise ValueError("pmats must be {}".format(want))
        self._ok(self._lib.phb_set_pmatrices(self._ctx, dptr(pmats)))

    def get_pmatrix(self, row, child):
        out = np.empty((self.n_cat, self.n_states, self.n_states))
        self._ok(self._lib.phb_get_pmatrix(self._ctx, int(row), int(child), dptr(out)))
        return out

    # ---- hot path ---------------------------------------------------------------------------
    def compute_partials(self, mode=_lib.PHB_MODE_AUTO):
        self._ok(self._lib.phb_compute_partials(self._ctx, int(mode)))

    def root_lnl(self, node_a, node_b, length, want_pattern=False, want_cat=False, root_pmats=None):
        total = ctypes.c_double(0.0)
        pattern = np.empty(self.n_patterns) if want_pattern else None
        cat = np.empty((self.n_patterns, self.n_cat)) if want_cat else None
        rp = None
        if root_pmats is not None:
            rp = _f64(root_pmats)
            if rp.shape != (2, self.n_cat, self.n_states, self.n_states):
                raise ValueError("root_pmats must be (2, K, A, A)")
        self._ok(self._lib.phb_root_lnl(self._ctx, int(node_a), int(node_b), float(length),
                                        dptr(rp) if rp is not None else None, ctypes.byref(total),
                                        dptr(pattern) if want_pattern else None, dptr(cat) if want_cat else None))
        return total.value, pattern, cat

    def lnl_resident(self, node_a, node_b, length, want_pattern=False):
        total = ctypes.c_double(0.0)
        pattern = np.empty(self.n_patterns) if want_pattern else None
        self._ok(self._lib.phb_lnl_resident(self._ctx, int(node_a), int(node_b), float(length), ctypes.byref(total),
                                            dptr(pattern) if want_pattern else None))
        return total.value, pattern

    @staticmethod
    def pack_codes(codes):
        """uint8 (n_tips, n_patterns) codes < 16 -> (n_tips, (n_patterns + 1) // 2) bytes, two patterns per byte
        (even pattern in the low nibble): the input format of ``lnl_from_host(..., packed=True)``."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if codes.ndim != 2:
            raise ValueError("codes must be (n_tips, n_patterns)")
        out = np.empty((codes.shape[0], (codes.shape[1] + 1) // 2), dtype=np.uint8)
        check(lib().phb_pack_codes(ctypes.c_void_p(codes.ctypes.data), codes.shape[0], codes.shape[1],
                                   ctypes.c_void_p(out.ctypes.data)))
        return out

    @staticmethod
    def split_codes(codes):
        """uint8 (n_tips, n_patterns) codes < 8 -> (low plane (n_tips, ceil(n/4)), high plane (n_tips, ceil(n/8))): 3 bits per
        code, the input format of ``lnl_from_host_split`` (look-up tables of at most 8 rows)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if codes.ndim != 2:
            raise ValueError("codes must be (n_tips, n_patterns)")
        low = np.empty((codes.shape[0], (codes.shape[1] + 3) // 4), dtype=np.uint8)
        high = np.empty((codes.shape[0], (codes.shape[1] + 7) // 8), dtype=np.uint8)
        check(lib().phb_split_codes(ctypes.c_void_p(codes.ctypes.data), codes.shape[0], codes.shape[1],
                                    ctypes.c_void_p(low.ctypes.data), ctypes.c_void_p(high.ctypes.data)))
        return low, high

    def _split_planes(self, planes):
        low, high = (np.ascontiguousarray(p, dtype=np.uint8) for p in planes)
        if low.shape != (self.n_tips, (self.n_patterns + 3) // 4) or high.shape != (self.n_tips, (self.n_patterns + 7) // 8):
            raise ValueError("split codes must be ({0}, ceil({1}/4)) and ({0}, ceil({1}/8))".format(self.n_tips, self.n_patterns))
        return low, high

    def lnl_from_host_split(self, planes, node_a, node_b, length, n_chunks=0, want_pattern=False):
        """``lnl_from_host`` from the two planes of ``split_codes`` (pinned host memory)."""
        low, high = self._split_planes(planes)
        total = ctypes.c_double(0.0)
        pattern = np.empty(self.n_patterns) if want_pattern else None
        self._ok(self._lib.phb_lnl_from_host_split(self._ctx, ctypes.c_void_p(low.ctypes.data), ctypes.c_void_p(high.ctypes.data),
                                                   int(n_chunks), int(node_a), int(node_b), float(length), ctypes.byref(total),
                                                   dptr(pattern) if want_pattern else None))
        return total.value, pattern

    def lnl_from_host_split_async(self, planes, node_a, node_b, length, n_chunks=0):
        low, high = self._split_planes(planes)
        self._keep["host_codes"] = (low, high)    # the copy engine reads them until the evaluation is complete
        self._ok(self._lib.phb_lnl_from_host_split_async(self._ctx, ctypes.c_void_p(low.ctypes.data), ctypes.c_void_p(high.ctypes.data),
                                                         int(n_chunks), int(node_a), int(node_b), float(length)))

    def lnl_from_host(self, codes, node_a, node_b, length, n_chunks=0, want_pattern=False, packed=False):
        """Evaluate starting from host tip codes (numpy uint8, ideally pinned); copy and compute overlap.
        codes is (n_tips, n_patterns), or with packed=True the output of ``pack_codes``."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        want = (self.n_tips, (self.n_patterns + 1) // 2) if packed else (self.n_tips, self.n_patterns)
        if codes.shape != want:
            raise ValueError("codes must be {}".format(want))
        total = ctypes.c_double(0.0)
        pattern = np.empty(self.n_patterns) if want_pattern else None
        fn = self._lib.phb_lnl_from_host_packed if packed else self._lib.phb_lnl_from_host
        self._ok(fn(self._ctx, ctypes.c_void_p(codes.ctypes.data), int(n_chunks), int(node_a), int(node_b),
                    float(length), ctypes.byref(total), dptr(pattern) if want_pattern else None))
        return total.value, pattern

    # ---- re-rooting in place -----------------------------------------------------------------
    def update_node(self, node, child_a, len_a, child_b, len_b):
        self._ok(self._lib.phb_update_node(self._ctx, int(node), int(child_a), float(len_a), int(child_b), float(len_b)))

    def branch_derivatives(self, node_a, node_b, lengths, chain_rule=True):
        lengths = _f64(np.atleast_1d(lengths))
        out = np.empty((lengths.shape[0], 3))
        self._ok(self._lib.phb_branch_derivatives(self._ctx, int(node_a), int(node_b), lengths.shape[0], dptr(lengths),
                                                  1 if chain_rule else 0, dptr(out)))
        return out

    # ---- stream-ordered forms: enqueue only, sums stay on the device (multi-GPU drivers) ---------------
    def lnl_resident_async(self, node_a, node_b, length):
        self._ok(self._lib.phb_lnl_resident_async(self._ctx, int(node_a), int(node_b), float(length)))

    def root_lnl_async(self, node_a, node_b, length):
        self._ok(self._lib.phb_root_lnl_async(self._ctx, int(node_a), int(node_b), float(length)))

    def lnl_from_host_packed_async(self, packed_codes, node_a, node_b, length, n_chunks=0):
        codes = np.ascontiguousarray(packed_codes, dtype=np.uint8)
        if codes.shape != (self.n_tips, (self.n_patterns + 1) // 2):
            raise ValueError("codes must be {}".format((self.n_tips, (self.n_patterns + 1) // 2)))
        self._keep["host_codes"] = codes          # the copy engine reads it until the evaluation is complete
        self._ok(self._lib.phb_lnl_from_host_packed_async(self._ctx, ctypes.c_void_p(codes.ctypes.data), int(n_chunks),
                                                          int(node_a), int(node_b), float(length)))

    def edge_derivatives_async(self, nodes, lengths, chain_rule=True):
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        lengths = _f64(lengths)
        if nodes.shape != lengths.shape or nodes.ndim != 1:
            raise ValueError("nodes and lengths must be 1-D and of equal length")
        self._ok(self._lib.phb_edge_derivatives_async(self._ctx, nodes.shape[0], iptr(nodes), dptr(lengths),
                                                      1 if chain_rule else 0))
        return nodes.shape[0]

    # ---- the scalar sum over the ranks of one box inside the reduction kernel ---------------------------
    PEER_HANDLE_BYTES = 64

    def peer_buffer(self):
        """Allocate this rank's exchange buffer (once) and return its CUDA IPC handle (``PEER_HANDLE_BYTES`` bytes)."""
        handle = ctypes.create_string_buffer(self.PEER_HANDLE_BYTES)
        self._ok(self._lib.phb_peer_buffer(self._ctx, handle))
        return handle.raw

    def peer_connect(self, rank, world, handles):
        """Map the exchange buffers of all ``world`` ranks (``handles``: their IPC handles concatenated in rank order)."""
        handles = bytes(handles)
        if len(handles) != self.PEER_HANDLE_BYTES * int(world):
            raise ValueError("need {} bytes of handles".format(self.PEER_HANDLE_BYTES * int(world)))
        self._ok(self._lib.phb_peer_connect(self._ctx, int(rank), int(world), handles))
        self.peer_world = int(world)

    def peer_sum_next(self):
        """The next stream-ordered scalar-lnL call leaves the sum over all connected ranks in the result buffer (every
        rank must make the same call)."""
        self._ok(self._lib.phb_peer_sum_next(self._ctx))

    @property
    def result_capacity(self):
        """Doubles the context's device result buffer holds (3 per edge of the tree, at least 256)."""
        ptr, cap = ctypes.c_void_p(), ctypes.c_int64(0)
        self._ok(self._lib.phb_device_result(self._ctx, ctypes.byref(ptr), ctypes.byref(cap)))
        return int(cap.value)

    # ---- pipelined host-fed evaluations: two in flight ---------------------------------------------------
    def host_fed_submit(self, codes, node_a, node_b, length, n_chunks=0):
        """Enqueue the evaluation of a new alignment given as pinned host codes - ``pack_codes`` output, or the
        ``(low, high)`` planes of ``split_codes`` - and return its slot (0 / 1).  Its copy runs under the walk of the
        evaluation submitted before; ``result_post(slot)`` + ``result_wait(slot)`` deliver the sum."""
        if isinstance(codes, (tuple, list)):
            low, high = self._split_planes(codes)
            hp = ctypes.c_void_p(high.ctypes.data)
        else:
            low = np.ascontiguousarray(codes, dtype=np.uint8)
            if low.shape != (self.n_tips, (self.n_patterns + 1) // 2):
                raise ValueError("packed codes must be {}".format((self.n_tips, (self.n_patterns + 1) // 2)))
            high, hp = None, None
        slot = ctypes.c_int(-1)
        self._ok(self._lib.phb_lnl_from_host_submit(self._ctx, ctypes.c_void_p(low.ctypes.data), hp, int(n_chunks), int(node_a),
                                                    int(node_b), float(length), ctypes.byref(slot)))
        self._keep["host_codes_slot{}".format(slot.value)] = (low, high)      # read by the copy engine until the walk has run
        return slot.value

    def result_post(self, slot):
        self._ok(self._lib.phb_result_post(self._ctx, int(slot)))

    def result_wait(self, slot):
        out = ctypes.c_double(0.0)
        self._ok(self._lib.phb_result_wait(self._ctx, int(slot), ctypes.byref(out)))
        return out.value

    def result_tensor(self, n, offset=0):
        """``n`` doubles of the context's device result buffer (from ``offset``) as a torch tensor VIEW (no copy): what a
        collective on the same stream reduces in place.  Needs the torch-owned workspace."""
        ws = self._keep.get("workspace")
        if ws is None:
            raise RuntimeError("result_tensor needs a torch-owned workspace (use_torch=True on a CUDA box)")
        ptr, cap = ctypes.c_void_p(), ctypes.c_int64(0)
        self._ok(self._lib.phb_device_result(self._ctx, ctypes.byref(ptr), ctypes.byref(cap)))
        if not 0 <= n + offset <= cap.value:
            raise ValueError("the device result buffer holds {} doubles".format(cap.value))
        import torch
        off = ptr.value - ws.data_ptr() + 8 * int(offset)
        return ws[off:off + 8 * int(n)].view(torch.float64)

    def result_fetch(self, n):
        out = np.empty(int(n))
        self._ok(self._lib.phb_result_fetch(self._ctx, int(n), dptr(out)))
        return out

    # ---- read-back --------------------------------------------------------------------------
    def get_partials(self, node):
        out = np.empty((self.n_patterns, self.n_cat, self.n_states))
        self._ok(self._lib.phb_get_partials(self._ctx, int(node), dptr(out)))
        return out

    def get_scalers(self, node):
        out = np.empty((self.n_patterns, self.n_cat))
        self._ok(self._lib.phb_get_scalers(self._ctx, int(node), dptr(out)))
        return out

    def get_root_partials(self):
        part = np.empty((self.n_patterns, self.n_cat, self.n_states))
        scal = np.empty((self.n_patterns, self.n_cat))
        self._ok(self._lib.phb_get_root_partials(self._ctx, dptr(part), dptr(scal)))
        return part, scal

    # ---- derivatives ------------------------------------------------------------------------
    def compute_up_partials(self, node_a, node_b, length):
        self._ok(self._lib.phb_compute_up_partials(self._ctx, int(node_a), int(node_b), float(length)))

    def edge_derivatives(self, nodes, lengths, chain_rule=True):
        nodes = np.ascontiguousarray(nodes, dtype=np.int32)
        lengths = _f64(lengths)
        if nodes.shape != lengths.shape or nodes.ndim != 1:
            raise ValueError("nodes and lengths must be 1-D and of equal length")
        out = np.empty((nodes.shape[0], 3))
        self._ok(self._lib.phb_edge_derivatives(self._ctx, nodes.shape[0], iptr(nodes), dptr(lengths),
                                                1 if chain_rule else 0, dptr(out)))
        return out
