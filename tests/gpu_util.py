"""Helpers for the GPU parity tests: run the same inputs through the CUDA C ABI and the CPU oracle."""
import numpy as np

from fmcmc_b200 import _abi as A


def kernel_kf(spec):
    k = spec["k"]
    fixed = np.broadcast_to(np.asarray(spec.get("fixed", False), dtype=bool), (k,))
    return int((~fixed).sum())


def kdraw_for(spec):
    kf = kernel_kf(spec)
    if spec["type"] in (A.KERNEL_ADAPT, A.KERNEL_RAM):
        return kf
    return kf if spec.get("scheme", A.SCHEME_JOINT) == A.SCHEME_JOINT else 1


def stream_kind(spec):
    t = spec["type"]
    if t in (A.KERNEL_UNIF, A.KERNEL_UNIF_REFLECTIVE, A.KERNEL_UMIRROR):
        return "unif"
    if t == A.KERNEL_RAM:
        return "t"
    return "normal"


def run_both(oracle, family, spec, initial, T, C, rng=None, path=0, burnin=0, thin=1, philox_seed=None,
             bulks=1, chain_offset=0):
    """Returns (cuda_out, oracle_out) lists of per-bulk dicts; kernel state is carried across bulks."""
    from fmcmc_b200.device import DeviceModel
    from gpu_util import kdraw_for, stream_kind  # noqa
    k = spec["k"]
    kf = kernel_kf(spec)
    dlen = A.state_len(spec["type"], k, kf)
    model = DeviceModel(family)
    if path:
        model.set_path(path)
    desc = family.marshal()
    ist_g = np.zeros((C, A.ISTATE_LEN), dtype=np.int64)
    dst_g = np.zeros((C, max(dlen, 1)))
    ist_o, dst_o = ist_g.copy(), dst_g.copy()
    init_g = init_o = np.ascontiguousarray(np.broadcast_to(np.asarray(initial, dtype=np.float64), (C, k)))
    outs_g, outs_o = [], []
    try:
        for b in range(bulks):
            if philox_seed is None:
                kd = kdraw_for(spec)
                kind = stream_kind(spec)
                logu = np.log(rng.random((C, T)))
                if kind == "normal":
                    z = rng.standard_normal((C, T, kd))
                elif kind == "unif":
                    z = rng.random((C, T, kd))
                else:
                    z = rng.standard_t(kf, size=(C, T, kd))
                mk = lambda: A.marshal_stream(A.STREAM_FED, logu=logu, z=z)
            else:
                mk = lambda: A.marshal_stream(A.STREAM_PHILOX, seed=philox_seed, run_index=b)
            g = model.run(spec, T, C, initial=init_g if b == 0 else None, burnin=burnin if b == 0 else 0,
                          thin=thin, stream=mk(), istate=ist_g, dstate=dst_g if dlen else None,
                          chain_offset=chain_offset)
            o = oracle.run(desc, spec, init_o, T, nchains=C, burnin=burnin if b == 0 else 0, thin=thin,
                           stream=mk(), istate=ist_o, dstate=dst_o, chain_offset=chain_offset, threads=8)
            init_o = o["ans"][:, -1, :]
            outs_g.append(g)
            outs_o.append(o)
    finally:
        model.close()
    return outs_g, outs_o, (ist_g, dst_g, ist_o, dst_o)


def col_rel_err(a, b):
    """Norm-wise relative error per parameter column: max|a-b| / max|b| over the whole run.  (An
    element-wise ratio is meaningless for samples that happen to pass near zero: their absolute
    rounding error is set by the O(1) operands they were summed from.)"""
    fin = np.isfinite(b)
    d = np.where(fin, np.abs(np.where(fin, a, 0.0) - np.where(fin, b, 0.0)), 0.0)
    bb = np.where(fin, np.abs(b), 0.0)
    if a.ndim == 3:
        return float(np.max(d.max(axis=(0, 1)) / np.maximum(bb.max(axis=(0, 1)), 1e-300)))
    return float(d.max() / max(bb.max(), 1e-300))


def elem_rel_err(a, b):
    """Element-wise relative error |a - b| / max(|b|, 1e-3 * column scale): the literal reading of "every sample
    within 1e-12 relative", with a floor so that samples passing near zero do not divide rounding noise by ~0."""
    fin = np.isfinite(b)
    d = np.where(fin, np.abs(np.where(fin, a, 0.0) - np.where(fin, b, 0.0)), 0.0)
    bb = np.where(fin, np.abs(b), 0.0)
    scale = bb.max(axis=(0, 1), keepdims=True) if a.ndim == 3 else bb.max()
    return float(np.max(d / np.maximum(np.maximum(bb, 1e-3 * scale), 1e-300)))


PARITY_LOG = []      # (what, name, norm-wise, element-wise) of every assert_parity call (printed by conftest at exit)


def assert_parity(g, o, rtol=1e-12, what="", elem_rtol=None):
    """Fed-stream contract: every accept/reject decision identical, samples within rtol norm-wise (relative to
    the parameter's scale, see col_rel_err) AND within elem_rtol element-wise (default 1000 rtol = what the norm-wise
    bound implies for an element at elem_rel_err's floor of 1e-3 of its column's scale; the run summary prints the
    worst figure actually seen, which for most kernels is the norm-wise one: the large elements carry the error)."""
    acc_g = np.any(g["ans"][:, 1:, :] != g["ans"][:, :-1, :], axis=2)
    acc_o = np.any(o["ans"][:, 1:, :] != o["ans"][:, :-1, :], axis=2)
    assert np.array_equal(acc_g, acc_o), f"{what}: accept/reject decisions differ at {np.argwhere(acc_g != acc_o)[:5]}"
    if elem_rtol is None:
        elem_rtol = 1000 * rtol
    for name in ("ans", "draws", "logpost"):
        a, b = g[name], o[name]
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin), f"{what}: {name} finiteness differs"
        assert np.array_equal(a[~fin], b[~fin], equal_nan=True), f"{what}: {name} non-finite values differ"
        err = col_rel_err(a, b)
        eerr = elem_rel_err(a, b)
        PARITY_LOG.append((what, name, err, eerr))
        assert err <= rtol, f"{what}: {name} max rel err {err:.3e} > {rtol}"
        assert eerr <= elem_rtol, f"{what}: {name} max element-wise rel err {eerr:.3e} > {elem_rtol}"
