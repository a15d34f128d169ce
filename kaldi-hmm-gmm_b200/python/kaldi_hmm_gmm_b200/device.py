"""Handle objects over the C ABI (include/khg_b200.h): a packed AmDiagGmm on the
device (`DeviceModel`) and its packed AccumAmDiagGmm statistics (`DeviceStats`).

Buffers are numpy arrays (host; copied inside the call) or torch CUDA tensors
(device; used in place).  All arithmetic happens in libkhg_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _cabi as A


class DeviceModel:
    """Packed, device-resident AmDiagGmm (reference csrc/am-diag-gmm.h:96: a vector of
    DiagGmm*, here one contiguous pack; pdf p owns Gaussians [offsets[p], offsets[p+1]))."""

    def __init__(self, dim: int, offsets):
        self.offsets = np.ascontiguousarray(offsets, np.int32)
        self.dim = int(dim)
        self.num_pdfs = self.offsets.size - 1
        self.num_gauss = int(self.offsets[-1])
        h = C.c_void_p()
        A.check(A.lib().khg_model_create(self.dim, self.num_pdfs, self.offsets.ctypes.data, C.byref(h)))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            try:
                A.lib().khg_model_destroy(self._h)
            except TypeError:  # interpreter shutdown: module globals already cleared
                pass
            self._h = None

    __del__ = close

    def upload(self, weights, means_invvars, inv_vars, gconsts=None) -> int:
        """Returns num_bad like DiagGmm::ComputeGconsts (csrc/diag-gmm.cc:103-147) when
        gconsts is None (computed on the device)."""
        miv = np.ascontiguousarray(means_invvars, np.float32)
        iv = np.ascontiguousarray(inv_vars, np.float32)
        assert miv.shape == (self.num_gauss, self.dim) and iv.shape == miv.shape
        w = None if weights is None else np.ascontiguousarray(weights, np.float32)
        gc = None if gconsts is None else np.ascontiguousarray(gconsts, np.float32)
        nb = C.c_int32(0)
        A.check(A.lib().khg_model_upload(self._h, None if w is None else w.ctypes.data, miv.ctypes.data, iv.ctypes.data,
                                         None if gc is None else gc.ctypes.data, C.byref(nb)))
        return nb.value

    def gconsts(self) -> np.ndarray:
        out = np.empty(self.num_gauss, np.float32)
        A.check(A.lib().khg_model_get_gconsts(self._h, out.ctypes.data))
        return out

    @classmethod
    def _from_handle(cls, handle) -> "DeviceModel":
        self = cls.__new__(cls)
        self._h = handle
        d, p, g = C.c_int32(), C.c_int32(), C.c_int32()
        A.check(A.lib().khg_model_info(handle, C.byref(d), C.byref(p), C.byref(g)))
        self.dim, self.num_pdfs, self.num_gauss = d.value, p.value, g.value
        self.offsets = np.empty(self.num_pdfs + 1, np.int32)
        A.check(A.lib().khg_model_download(handle, self.offsets.ctypes.data, None, None, None, None))
        return self

    def download(self):
        """Host copies of the packed parameters: dict(offsets, weights, means_invvars, inv_vars, gconsts)."""
        G, D = self.num_gauss, self.dim
        w, gc = np.empty(G, np.float32), np.empty(G, np.float32)
        miv, iv = np.empty((G, D), np.float32), np.empty((G, D), np.float32)
        offs = np.empty(self.num_pdfs + 1, np.int32)
        A.check(A.lib().khg_model_download(self._h, offs.ctypes.data, w.ctypes.data, miv.ctypes.data, iv.ctypes.data, gc.ctypes.data))
        return dict(offsets=offs, weights=w, means_invvars=miv, inv_vars=iv, gconsts=gc)

    # pickle / torch.save round trip of the packed model (the reference pickles
    # (weights, inv_vars, means_invvars) per pdf, python/csrc/am-diag-gmm.cc:47-71): the state is the
    # packed host copy, the device pack is rebuilt on load with the stored gconsts.
    def __getstate__(self):
        st = self.download()
        st["dim"] = self.dim
        return st

    def __setstate__(self, st):
        self.__init__(st["dim"], st["offsets"])
        self.upload(st["weights"], st["means_invvars"], st["inv_vars"], st["gconsts"])

    def dense_kernel(self) -> int:
        """1 = SIMT, 2 = tcgen05 3xTF32, 3 = tcgen05 3xFP16 (what an in-range call runs)."""
        k = C.c_int32()
        A.check(A.lib().khg_model_dense_kernel(self._h, C.byref(k)))
        return k.value

    def stats_kernel(self) -> int:
        """1 = fp32 kernel, 3 = tcgen05 kernel (fp16 hi/lo split tile): what the bucketed statistics pass runs."""
        k = C.c_int32()
        A.check(A.lib().khg_model_stats_kernel(self._h, C.byref(k)))
        return k.value

    def set_kernel(self, kernel: int):
        A.check(A.lib().khg_model_set_kernel(self._h, kernel))

    def set_stream(self, cuda_stream: int):
        A.check(A.lib().khg_model_set_stream(self._h, cuda_stream))

    def sync(self):
        A.check(A.lib().khg_model_sync(self._h))

    # -- likelihoods ------------------------------------------------------
    def loglikes_all_pdfs(self, feats, scale: float = 1.0, layout: int = A.KHG_FRAME_MAJOR, out=None):
        """(T,P) [frame-major] or (P,T) [pdf-major] block of per-pdf log-likelihoods."""
        fp, floc = A.ptr(feats, np.float32)
        T = int(feats.shape[0])
        assert feats.shape[1] == self.dim, "Dim mismatch"
        shape = (T, self.num_pdfs) if layout == A.KHG_FRAME_MAJOR else (self.num_pdfs, T)
        if out is None:
            if floc == A.KHG_DEVICE:
                import torch

                out = torch.empty(shape, dtype=torch.float32, device=feats.device)
            else:
                out = np.empty(shape, np.float32)
        op, oloc = A.ptr(out, np.float32)
        ld = int(out.shape[1]) if len(out.shape) == 2 else shape[1]
        A.check(A.lib().khg_loglikes_all_pdfs(self._h, fp, T, floc, scale, layout, op, ld, oloc))
        return out

    def loglikes_pdf_subset(self, feats, pdf_subset, scale: float = 1.0):
        """(len(pdf_subset), T) pdf-major block of the listed pdfs only (host or device buffers)."""
        fp, floc = A.ptr(feats, np.float32)
        T = int(feats.shape[0])
        sub = np.ascontiguousarray(pdf_subset, np.int32)
        if floc == A.KHG_DEVICE:
            import torch

            out = torch.empty((sub.size, T), dtype=torch.float32, device=feats.device)
        else:
            out = np.empty((sub.size, T), np.float32)
        op, oloc = A.ptr(out, np.float32)
        A.check(A.lib().khg_loglikes_pdf_subset(self._h, fp, T, floc, sub.ctypes.data, sub.size, scale, op, T, oloc))
        return out

    def split_by_count(self, state_occs, target_components: int, perturb_factor: float = 0.01, power: float = 0.2,
                       min_count: float = 20.0, randn=None, seed: int = 0) -> "DeviceModel":
        """AmDiagGmm::SplitByCount (reference csrc/am-diag-gmm.cc:72-89; defaults of scripts/gmm_est.py) on the
        device: a NEW DeviceModel with the mixed-up Gaussians.  randn (rows x dim): the standard-normal
        vector of every split in the reference's order, or None to draw them from `seed`."""
        occ = np.ascontiguousarray(state_occs, np.float32)
        assert occ.size == self.num_pdfs
        rn = None if randn is None else np.ascontiguousarray(randn, np.float32).reshape(-1, self.dim)
        nh, ng = C.c_void_p(), C.c_int32()
        A.check(A.lib().khg_model_split_by_count(self._h, occ.ctypes.data, int(target_components), perturb_factor, power,
                                                 min_count, None if rn is None else rn.ctypes.data,
                                                 0 if rn is None else rn.shape[0], seed, C.byref(nh), C.byref(ng)))
        return DeviceModel._from_handle(nh)

    def merge_by_count(self, state_occs, target_components: int, power: float = 0.2, min_count: float = 20.0) -> "DeviceModel":
        """AmDiagGmm::MergeByCount (reference csrc/am-diag-gmm.cc:91-108) on the device: a NEW DeviceModel."""
        occ = np.ascontiguousarray(state_occs, np.float32)
        assert occ.size == self.num_pdfs
        nh, ng = C.c_void_p(), C.c_int32()
        A.check(A.lib().khg_model_merge_by_count(self._h, occ.ctypes.data, int(target_components), power, min_count,
                                                 C.byref(nh), C.byref(ng)))
        return DeviceModel._from_handle(nh)

    def gaussian_selection(self, pdf: int, feats, num_gselect: int, preselect=None, want_loglikes: bool = False):
        """DiagGmm::GaussianSelection / GaussianSelectionPreselect (reference csrc/diag-gmm.cc:202-366) of
        pdf `pdf` for all rows of feats: (total log-like, indices int32 [T, k], per-frame log-like [T]
        [, log-likes of the selected components [T, k]])."""
        fp, floc = A.ptr(feats, np.float32)
        T = int(feats.shape[0])
        ng = int(self.offsets[pdf + 1] - self.offsets[pdf])
        pre = None if preselect is None else np.ascontiguousarray(preselect, np.int32)
        n = ng if pre is None else pre.size
        k = min(int(num_gselect), n)
        idx = np.empty((T, k), np.int32)
        ll = np.empty((T, k), np.float32) if want_loglikes else None
        fl = np.empty(T, np.float32)
        tot = C.c_double(0.0)
        A.check(A.lib().khg_gaussian_selection(self._h, pdf, fp, T, floc, None if pre is None else pre.ctypes.data,
                                               0 if pre is None else pre.size, int(num_gselect), idx.ctypes.data,
                                               None if ll is None else ll.ctypes.data, fl.ctypes.data, C.byref(tot)))
        return (tot.value, idx, fl, ll) if want_loglikes else (tot.value, idx, fl)

    def pdf_loglikes(self, pdf: int, feats: np.ndarray) -> np.ndarray:
        feats = np.ascontiguousarray(np.atleast_2d(feats), np.float32)
        if feats.shape[1] != self.dim:
            raise RuntimeError(f"DiagGmm::LogLikelihoods, dimension mismatch {feats.shape[1]} vs. {self.dim}")
        ng = int(self.offsets[pdf + 1] - self.offsets[pdf])
        out = np.empty((feats.shape[0], ng), np.float32)
        A.check(A.lib().khg_pdf_loglikes(self._h, pdf, feats.ctypes.data, feats.shape[0], A.KHG_HOST, out.ctypes.data))
        return out

    def pdf_posteriors(self, pdf: int, feats: np.ndarray, want_post: bool = True):
        feats = np.ascontiguousarray(np.atleast_2d(feats), np.float32)
        if feats.shape[1] != self.dim:
            raise RuntimeError(f"DiagGmm::LogLikelihoods, dimension mismatch {feats.shape[1]} vs. {self.dim}")
        ng = int(self.offsets[pdf + 1] - self.offsets[pdf])
        post = np.empty((feats.shape[0], ng), np.float32) if want_post else None
        ll = np.empty(feats.shape[0], np.float32)
        A.check(A.lib().khg_pdf_posteriors(self._h, pdf, feats.ctypes.data, feats.shape[0], A.KHG_HOST,
                                           None if post is None else post.ctypes.data, ll.ctypes.data))
        return ll, post


class GraphBatch:
    """Compiled training graphs of a batch of utterances as plain arrays (khg_graph_batch,
    include/khg_b200.h): what a maintainer exports from `fst::VectorFst<StdArc>` after
    AddTransitionProbs (reference scripts/gmm_align_compiled.py:36-41).

    graphs: sequence of objects with arc_offsets[S+1], ilabel[A], olabel[A] (optional), weight[A],
    nextstate[A], final[S] (inf = not final), start — arcs sorted by source state.
    num_frames: frames of each utterance (rows of the concatenated feature matrix)."""

    def __init__(self, graphs, num_frames):
        assert len(graphs) == len(num_frames)
        i32 = lambda x: np.ascontiguousarray(x, np.int32)  # noqa: E731
        self.n_utts = len(graphs)
        self.frame_offsets = np.concatenate([[0], np.cumsum(np.asarray(num_frames, np.int64))]).astype(np.int64)
        ns = [int(np.asarray(g.arc_offsets).size) - 1 for g in graphs]
        self.state_offsets = np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)
        offs, base = [np.zeros(1, np.int64)], 0
        for g in graphs:
            ao = np.asarray(g.arc_offsets, np.int64)
            offs.append(ao[1:] + base)
            base += int(ao[-1])
        self.arc_offsets = i32(np.concatenate(offs))
        cat = lambda name, dt: np.ascontiguousarray(  # noqa: E731
            np.concatenate([np.asarray(getattr(g, name), dt) for g in graphs]) if graphs else np.zeros(0, dt), dt)
        self.arc_ilabel, self.arc_nextstate = cat("ilabel", np.int32), cat("nextstate", np.int32)
        self.arc_weight, self.final_cost = cat("weight", np.float32), cat("final", np.float32)
        self.arc_olabel = cat("olabel", np.int32) if all(hasattr(g, "olabel") for g in graphs) else None
        self.start_state = i32([g.start for g in graphs])

    def c_struct(self) -> "A.GraphBatch":
        d = lambda a: a.ctypes.data  # noqa: E731
        return A.GraphBatch(self.n_utts, d(self.frame_offsets), d(self.state_offsets), d(self.arc_offsets), d(self.arc_ilabel),
                            d(self.arc_nextstate), d(self.arc_weight), d(self.start_state), d(self.final_cost))


def align_batch(model: DeviceModel, graphs: GraphBatch, feats, tid2pdf, acoustic_scale: float = 1.0, beam: float = 200.0,
                retry_beam: float = 0.0, want_paths: bool = True, pdf_ids_out=None):
    """gmm-align-compiled for a batch of utterances in ONE call (khg_align_batch): dense
    log-likelihoods (K1) + per-utterance Viterbi on the device.  Returns dict(alignment int32
    per frame, status per utterance (0 ok / 1 ok after retry / 2 failed), like per utterance,
    path_offsets, path_arcs (absolute arc ids, epsilons included), words (if the graphs carry olabels))."""
    fp, floc = A.ptr(feats, np.float32)
    T = int(graphs.frame_offsets[-1])
    assert int(feats.shape[0]) >= T and int(feats.shape[1]) == model.dim
    t2p = np.ascontiguousarray(tid2pdf, np.int32)
    U = graphs.n_utts
    ali = np.zeros(T, np.int32)
    status = np.zeros(U, np.int32)
    like = np.zeros(U, np.float32)
    poff = np.zeros(U + 1, np.int64)
    cap = int(T + T // 2 + 16 * U + 1024) if want_paths else 0
    # a path has one arc per frame plus its epsilon arcs: T + (epsilon arcs on the path); retried below if short
    pdp, ploc = A.ptr(pdf_ids_out, np.int32)
    assert pdf_ids_out is None or ploc == A.KHG_DEVICE
    gs = graphs.c_struct()
    while True:
        paths = np.zeros(cap, np.int32) if want_paths else None
        st = A.lib().khg_align_batch(model._h, C.byref(gs), fp, floc, t2p.ctypes.data, t2p.size, acoustic_scale, beam, retry_beam,
                                     ali.ctypes.data, status.ctypes.data, like.ctypes.data,
                                     None if paths is None else paths.ctypes.data, poff.ctypes.data, cap, pdp)
        if st != A.KHG_OK and want_paths and b"path_capacity" in (A.lib().khg_last_error() or b""):
            cap *= 4
            continue
        A.check(st)
        break
    out = dict(alignment=ali, status=status, like=like, path_offsets=poff, path_arcs=None if paths is None else paths[:poff[-1]])
    if want_paths and graphs.arc_olabel is not None:
        ol = graphs.arc_olabel[out["path_arcs"]]
        out["words"] = [ol[poff[u]:poff[u + 1]][ol[poff[u]:poff[u + 1]] != 0] for u in range(U)]
    return out


def align_utterance_host(graphs: GraphBatch, utt: int, loglikes, tid2row, acoustic_scale: float = 1.0, beam: float = 200.0,
                         retry_beam: float = 0.0):
    """The reference's FasterDecoder + AlignUtteranceWrapper on the host for utterance `utt`, on a (rows, T) block
    of SCALED log-likelihoods (khg_align_utterance_host; needs no GPU).  Returns dict(status, alignment, like, path)."""
    ll = np.ascontiguousarray(loglikes, np.float32)
    t2r = np.ascontiguousarray(tid2row, np.int32)
    T = int(graphs.frame_offsets[utt + 1] - graphs.frame_offsets[utt])
    assert ll.ndim == 2 and ll.shape[1] >= T
    ali = np.zeros(T, np.int32)
    n_arcs_u = int(graphs.arc_offsets[graphs.state_offsets[utt + 1]] - graphs.arc_offsets[graphs.state_offsets[utt]])
    cap = T + n_arcs_u + 16
    path = np.zeros(cap, np.int32)
    status, like, plen = C.c_int32(), C.c_float(), C.c_int32()
    gs = graphs.c_struct()
    A.check(A.lib().khg_align_utterance_host(C.byref(gs), utt, ll.ctypes.data, ll.shape[1], t2r.ctypes.data, t2r.size, acoustic_scale,
                                             beam, retry_beam, ali.ctypes.data, C.byref(status), C.byref(like), path.ctypes.data, cap,
                                             C.byref(plen)))
    return dict(status=status.value, alignment=ali, like=like.value, path=path[:plen.value])


class DeviceStats:
    """Packed device AccumAmDiagGmm: [occ G | mean G*D | var G*D | tot_like, tot_frames] fp64."""

    def __init__(self, model: DeviceModel, flags: int = 0xF):
        self.model = model
        h = C.c_void_p()
        A.check(A.lib().khg_stats_create(model._h, flags, C.byref(h)))
        self._h = h
        f = C.c_uint16()
        A.check(A.lib().khg_stats_flags(h, C.byref(f)))
        self.flags = f.value

    def close(self):
        if getattr(self, "_h", None):
            try:
                A.lib().khg_stats_destroy(self._h)
            except TypeError:  # interpreter shutdown: module globals already cleared
                pass
            self._h = None

    __del__ = close

    def zero(self):
        A.check(A.lib().khg_stats_zero(self._h))

    def device_buffer(self):
        """(device pointer, number of doubles) of the packed stats — for the NCCL all-reduce."""
        p, n = C.c_void_p(), C.c_int64()
        A.check(A.lib().khg_stats_device_buffer(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def as_torch(self):
        """Zero-copy torch.float64 CUDA view of the packed buffer."""
        import torch

        p, n = self.device_buffer()

        class _Holder:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (p, False), "version": 3, "strides": None}

        return torch.as_tensor(_Holder(), device="cuda")

    def download(self):
        G, D = self.model.num_gauss, self.model.dim
        occ = np.empty(G, np.float64)
        mean = np.empty((G, D), np.float64) if self.flags & 1 else None
        var = np.empty((G, D), np.float64) if self.flags & 2 else None
        tot = np.empty(2, np.float64)
        A.check(A.lib().khg_stats_download(self._h, occ.ctypes.data, None if mean is None else mean.ctypes.data,
                                           None if var is None else var.ctypes.data, tot.ctypes.data))
        return dict(occ=occ, mean=mean, var=var, tot_like=float(tot[0]), tot_frames=float(tot[1]))

    def upload(self, occ=None, mean=None, var=None, totals=None):
        def p(a):
            return None if a is None else np.ascontiguousarray(a, np.float64)

        occ, mean, var, totals = p(occ), p(mean), p(var), p(totals)
        A.check(A.lib().khg_stats_upload(self._h, *(None if a is None else a.ctypes.data for a in (occ, mean, var, totals))))

    def add(self, scale: float, other: "DeviceStats"):
        A.check(A.lib().khg_stats_add(self._h, scale, other._h))

    def scale(self, scale: float):
        A.check(A.lib().khg_stats_scale(self._h, scale))

    def mle_update(self, update_flags: int = 0xF, min_gaussian_weight: float = 1e-5, min_gaussian_occupancy: float = 10.0,
                   min_variance: float = 0.001, remove_low_count_gaussians: bool = True):
        """Device M-step (MleAmDiagGmmUpdate, reference csrc/mle-am-diag-gmm.cc:153-202).
        Returns (new DeviceModel, dict(obj_change, count, floored_elements, floored_gaussians, removed_gaussians))."""

        class _Opts(C.Structure):
            _fields_ = [("w", C.c_float), ("occ", C.c_float), ("var", C.c_double), ("rm", C.c_int32)]

        o = _Opts(min_gaussian_weight, min_gaussian_occupancy, min_variance, int(remove_low_count_gaussians))
        nh = C.c_void_p()
        oc, cnt = C.c_float(), C.c_float()
        fe, fg, rg = C.c_int32(), C.c_int32(), C.c_int32()
        A.check(A.lib().khg_mle_update(self.model._h, self._h, C.byref(o), update_flags, C.byref(nh), C.byref(oc), C.byref(cnt),
                                       C.byref(fe), C.byref(fg), C.byref(rg)))
        return DeviceModel._from_handle(nh), dict(obj_change=oc.value, count=cnt.value, floored_elements=fe.value,
                                                  floored_gaussians=fg.value, removed_gaussians=rg.value)

    # -- accumulation -----------------------------------------------------
    def acc_stats_ali(self, feats, pdf_ids, frame_weights=None, per_frame=None, want_total: bool = True) -> Optional[float]:
        """gmm-acc-stats-ali for T frames (reference scripts/gmm_acc_stats_ali.py:46-56)."""
        fp, l0 = A.ptr(feats, np.float32)
        ip, l1 = A.ptr(pdf_ids, np.int32)
        wp, l2 = A.ptr(frame_weights, np.float32)
        pp, l3 = A.ptr(per_frame, np.float32)
        loc = A.same_loc(l0, l1, None if frame_weights is None else l2, None if per_frame is None else l3)
        tot = C.c_double(0.0)
        A.check(A.lib().khg_acc_stats_ali(self.model._h, self._h, fp, int(feats.shape[0]), loc, ip, wp, pp,
                                          C.byref(tot) if want_total else None))
        return tot.value if want_total else None

    def acc_stats_ali_tids(self, feats: np.ndarray, tids, tid2pdf, trans_accs: Optional[np.ndarray] = None) -> float:
        feats = np.ascontiguousarray(feats, np.float32)
        tids = np.ascontiguousarray(tids, np.int32)
        tid2pdf = np.ascontiguousarray(tid2pdf, np.int32)
        if trans_accs is not None:
            assert trans_accs.dtype == np.float64 and trans_accs.size == tid2pdf.size
        tot = C.c_double(0.0)
        A.check(A.lib().khg_acc_stats_ali_tids(self.model._h, self._h, feats.ctypes.data, feats.shape[0], tids.ctypes.data,
                                               tid2pdf.ctypes.data, tid2pdf.size - 1,
                                               None if trans_accs is None else trans_accs.ctypes.data, C.byref(tot)))
        return tot.value

    def acc_from_posteriors(self, pdf: int, feats: np.ndarray, post: np.ndarray):
        feats = np.ascontiguousarray(np.atleast_2d(feats), np.float32)
        post = np.ascontiguousarray(np.atleast_2d(post), np.float32)
        A.check(A.lib().khg_acc_from_posteriors(self.model._h, self._h, pdf, feats.ctypes.data, feats.shape[0], A.KHG_HOST, post.ctypes.data))

    def estep(self, feats, pdf_ids, loglikes_out, frame_weights=None, chunk_frames: int = 0, want_total: bool = False):
        """Dense all-pdf log-likelihoods (into the reused device block `loglikes_out`,
        pdf-major) + alignment statistics for one batch."""
        fp, l0 = A.ptr(feats, np.float32)
        ip, l1 = A.ptr(pdf_ids, np.int32)
        wp, l2 = A.ptr(frame_weights, np.float32)
        loc = A.same_loc(l0, l1, None if frame_weights is None else l2)
        op, lo = A.ptr(loglikes_out, np.float32)
        assert lo == A.KHG_DEVICE, "loglikes_out must be a device buffer"
        tot = C.c_double(0.0)
        A.check(A.lib().khg_estep(self.model._h, self._h, fp, int(feats.shape[0]), loc, ip, wp, op,
                                  int(loglikes_out.shape[1]), chunk_frames, C.byref(tot) if want_total else None))
        return tot.value if want_total else None
