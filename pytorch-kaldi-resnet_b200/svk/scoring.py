"""Trial scoring on the GPU: cosine scoring of trial pairs, cohort top-k statistics for adaptive s-norm, s-norm apply.

replaces the per-trial / per-utterance Python loops of the reference:
    cosine_score.py:60-65            -> cosine_scores()          (svk_cosine_score_pairs)
    compute_topk_mean_std.py:10-23   -> cohort_topk_meanstd()    (svk_l2norm_rows_fwd + svk_sgemm + svk_topk_meanstd)
    adaptive_snorm.py:28-38          -> snorm_apply()            (svk_snorm_apply)
    compute_mean.py:9-20             -> global_mean()            (svk_col_mean)
    compute_speaker_mean.py:9-30     -> speaker_means()          (svk_segment_mean)
    compute_eer.py:35-102, local/compute_min_dcf.py:54-120 -> det_metrics()  (svk_sort_pairs_f64 + svk_det_metrics)
Embeddings are float32 on the device, exactly what the reference feeds torch (`torch.FloatTensor(vec - mean)`).
"""
import torch

from . import lib as _lib
from . import ops  # noqa: F401  (registers torch.ops.svk.*)
from .lib import call


def _st():
    return torch.cuda.current_stream().cuda_stream


def _f32(t, device):
    return torch.as_tensor(t, dtype=torch.float32, device=device).contiguous()


def _i32(t, device):
    return torch.as_tensor(t, dtype=torch.int32, device=device).contiguous()


def cosine_scores(enroll, test, mean, idx_enroll, idx_test, device="cuda"):
    """score[t] = cosine(enroll[ie[t]] - mean, test[it[t]] - mean); mean may be None (already subtracted)."""
    E, T = _f32(enroll, device), _f32(test, device)
    ie, it = _i32(idx_enroll, device), _i32(idx_test, device)
    m = None if mean is None else _f32(mean, device)
    return torch.ops.svk.cosine_score_pairs(E, T, m, ie, it)


def l2_normalize_rows(x, eps=1e-12):
    x = x.contiguous()
    out = torch.empty_like(x)
    call.svk_l2norm_rows_fwd(x.data_ptr(), out.data_ptr(), 0, x.shape[0], x.shape[1], eps, _st())
    return out


SCORE_BLOCK_ROWS = 4096      # measured (profiles/r02_snorm_blocks.md): 256-row blocks keep the score block in L2 but lose more
                             # to launch / tail overhead (2.5 M rows/s) than 4,096-row blocks pay in HBM traffic (3.6 M rows/s)


def cohort_topk_meanstd(vecs, cohort, topk=300, block_rows=None, device="cuda", tf32=False):
    """For every row v of `vecs`: scores = normalize(cohort) @ normalize(v); top-k; (mean, unbiased std).
    Both inputs must already be mean-subtracted (compute_topk_mean_std.py:41-48).  Returns (mean, std) tensors.
    tf32=True computes the score matrix on the tensor cores (tcgen05 kind::tf32; scores within ~2e-4 of fp32, inside
    the 1e-3 north-star bound); the default keeps exact fp32 products like the reference.
    The full (n, cohort) score matrix never exists: rows are processed in blocks of `block_rows` (default SCORE_BLOCK_ROWS),
    written by the GEMM and read back twice by the select (thread-local top-2 -> threshold -> candidate list)."""
    X, Cm = _f32(vecs, device), _f32(cohort, device)
    n, D = X.shape
    nc = Cm.shape[0]
    if nc < topk:
        raise ValueError("cohort of %d rows is smaller than topk=%d (torch.topk would fail too)" % (nc, topk))
    if n == 0:
        return torch.empty(0, dtype=torch.float32, device=device), torch.empty(0, dtype=torch.float32, device=device)
    Cn = l2_normalize_rows(Cm)
    Xn = l2_normalize_rows(X)
    if block_rows is None:
        block_rows = SCORE_BLOCK_ROWS
    block_rows = max(1, min(block_rows, n))
    scores = torch.empty(block_rows, nc, dtype=torch.float32, device=device)
    st = _st()
    means, stds = [], []
    for lo in range(0, n, block_rows):
        rows = min(block_rows, n - lo)
        xb = Xn[lo:lo + rows]
        if tf32 and D % 4 == 0:
            from . import lib as _lib
            need = _lib.load().svk_gemm_tf32_workspace_bytes(rows, nc, D)
            ws = torch.empty(need // 4 + 4, dtype=torch.float32, device=device) if need else None
            call.svk_gemm_tf32(xb.data_ptr(), D, 1, Cn.data_ptr(), D, 1, scores.data_ptr(), nc, rows, nc, D, 0,
                               0 if ws is None else ws.data_ptr(), need, st)
        else:
            call.svk_sgemm(xb.data_ptr(), D, 1, Cn.data_ptr(), 1, D, scores.data_ptr(), nc, rows, nc, D, 1.0, 0.0, 0, st)
        m, sd = torch.ops.svk.topk_meanstd(scores[:rows], topk)
        means.append(m)
        stds.append(sd)
    return torch.cat(means), torch.cat(stds)


def snorm_apply(scores, idx_enroll, idx_test, mean_e, std_e, mean_t, std_t, device="cuda"):
    s = _f32(scores, device)
    ie, it = _i32(idx_enroll, device), _i32(idx_test, device)
    me, se, mt, sd = _f32(mean_e, device), _f32(std_e, device), _f32(mean_t, device), _f32(std_t, device)
    return torch.ops.svk.snorm_apply(s, ie, it, me, se, mt, sd)


def global_mean(vecs, device="cuda"):
    """Mean over all rows (compute_mean.py:17).  Returns a float32 tensor [D]."""
    X = _f32(vecs, device)
    n, D = X.shape
    out = torch.empty(D, dtype=torch.float32, device=device)
    ws = torch.empty(_lib.load().svk_col_mean_workspace_bytes(n, D), dtype=torch.uint8, device=device)
    call.svk_col_mean(X.data_ptr(), n, D, out.data_ptr(), ws.data_ptr(), ws.numel(), _st())
    return out


def speaker_means(vecs, seg_ids, n_seg, device="cuda"):
    """Per-speaker mean of the rows of `vecs` (row r belongs to speaker seg_ids[r]), accumulated in row order like
    compute_speaker_mean.py:16-27.  Returns a float32 tensor [n_seg, D]."""
    import numpy as np
    X = _f32(vecs, device)
    seg = np.asarray(seg_ids, dtype=np.int64)
    order = np.argsort(seg, kind="stable").astype(np.int32)              # rows grouped by speaker, file order inside
    counts = np.bincount(seg, minlength=n_seg)
    offsets = np.zeros(n_seg + 1, dtype=np.int32)
    np.cumsum(counts, out=offsets[1:])
    out = torch.empty(n_seg, X.shape[1], dtype=torch.float32, device=device)
    order_d, offsets_d = _i32(order, device), _i32(offsets, device)       # keep them alive across the launch
    call.svk_segment_mean(X.data_ptr(), order_d.data_ptr(), offsets_d.data_ptr(), n_seg, X.shape[1], out.data_ptr(), _st())
    return out


def sort_pairs(keys, vals, device="cuda"):
    """Stable ascending sort of float64 keys with int32 payloads on the device."""
    k = torch.as_tensor(keys, dtype=torch.float64, device=device).contiguous()
    v = _i32(vals, device)
    n = k.numel()
    ko, vo = torch.empty_like(k), torch.empty_like(v)
    ws = torch.empty(_lib.load().svk_sort_pairs_f64_workspace_bytes(n), dtype=torch.uint8, device=device)
    call.svk_sort_pairs_f64(k.data_ptr(), v.data_ptr(), ko.data_ptr(), vo.data_ptr(), n, ws.data_ptr(), ws.numel(), _st())
    return ko, vo


def det_metrics(scores, labels, p_target=0.01, c_miss=1.0, c_fa=1.0, device="cuda"):
    """EER and minDCF of a scored trial list (labels: 1 = target, 0 = non-target), computed like the reference:
    stable ascending sort of the float64 scores, thresholds = sorted scores, fnr / fpr from cumulative counts.
    Returns dict(eer, eer_index, min_dcf, min_dcf_threshold, n_target, n_nontarget)."""
    import numpy as np
    lab = np.asarray(labels, dtype=np.int32)
    n, n_t = lab.size, int(lab.sum())
    ks, ls = sort_pairs(scores, lab, device)
    out = torch.empty(6, dtype=torch.float64, device=device)
    ws = torch.empty(_lib.load().svk_det_metrics_workspace_bytes(n), dtype=torch.uint8, device=device)
    call.svk_det_metrics(ls.data_ptr(), n, n_t, float(p_target), float(c_miss), float(c_fa), out.data_ptr(), ws.data_ptr(),
                         ws.numel(), _st())
    o = out.cpu().numpy()
    c_def = min(c_miss * p_target, c_fa * (1 - p_target))                # compute_min_dcf.py:104
    return {"eer": float(o[0]), "eer_index": int(o[1]), "min_dcf": float(o[2]) / c_def,
            "min_dcf_threshold": float(ks[int(o[3])].item()), "n_target": int(o[4]), "n_nontarget": int(o[5])}
