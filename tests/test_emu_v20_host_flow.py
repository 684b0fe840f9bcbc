"""The HOST orchestration of scan variant 20 (caduceus_b200/functional.py: scan_fwd -> scan_fwd_segmented, scan_fixup with
seg_ctx) run on CPU: the library handle is replaced by a stand-in whose scan entry point goes to the SIMT emulation of the
kernel source (tests/emu/) and whose helper entry points (transpose, carry composition, carry fix-up) are restated in numpy.  What this pins is the Python
between the kernels — segment count, buffer shapes, argument blocks, the order of the launches, the shard flow
(want_state -> all_gather stand-in -> scan_fixup(seg_ctx)) — which otherwise runs for the first time on the GPU box.
TEST INFRASTRUCTURE ONLY: the product path loads the CUDA library and nothing else (tests/test_host.py checks that it fails
loudly without it)."""
import ctypes as C

import numpy as np
import pytest
import torch

from caduceus_b200 import _lib, functional as CF
from scan_boundary_ref import _problem, boundary_ref
from emu_build import emu  # noqa: F401  (module-scoped fixture: builds tests/emu/libemu_scan.so)
from scan_boundary_ref import _silu, _softplus

N = 16


def _arr(ptr, shape, dtype=np.float32):
    n = int(np.prod(shape))
    buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr if isinstance(ptr, int) else ptr.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


class _EmuLib:
    """Same entry points as libcaduceus_b200.so for what variant 20's host flow calls."""

    def __init__(self, emu_lib):
        self.emu, self.calls = emu_lib, []

    def cad_sm_count(self):
        return 148

    def cad_scan_chunk_len(self):
        return 512

    def cad_last_error(self):
        return b"emulated"

    def cad_bc_transpose(self, bc, bcT, njobs, N2, L, ldbc, stream):
        self.calls.append("transpose")
        Lp = (L + 255) // 256 * 256
        src, dst = _arr(bc, (njobs, N2, ldbc)), _arr(bcT, (njobs, Lp, N2))
        dst[:] = 0
        dst[:, :L] = src[:, :, :L].transpose(0, 2, 1)
        return 0

    def cad_seg_carry(self, st, ds, A2, pset, h0, carry, hlast, dtsum, njobs, nseg, E, stream):
        self.calls.append("carry")
        st, ds = _arr(st, (njobs, nseg, E, N)), _arr(ds, (njobs, nseg, E))
        ps = _arr(pset, (njobs,), np.int32)
        a2 = _arr(A2, (int(ps.max()) + 1, E, N))[ps]
        h = np.zeros((njobs, E, N), np.float32) if not h0 else _arr(h0, (njobs, E, N)).copy()
        for s in range(nseg):
            if carry:
                _arr(carry, (njobs, nseg, E, N))[:, s] = h
            h = (np.exp2(a2 * ds[:, s, :, None]) * h + st[:, s]).astype(np.float32)
        if hlast:
            _arr(hlast, (njobs, E, N))[:] = h
        if dtsum:
            _arr(dtsum, (njobs, E))[:] = ds.sum(1)
        return 0

    def cad_bimamba_scan_fwd(self, ref, stream):
        a = ref._obj
        self.calls.append(f"scan v{a.variant} nseg {a.nseg} W {a.channels_per_cta}")
        assert a.variant == 20 and not a.h0 and not a.hlast and not a.dtsum and not a.chunk_state
        return self.emu.emu_scan_v20(C.byref(a), a.channels_per_cta)

    def cad_bimamba_scan_fixup(self, ref, stream):
        """float64 restatement of the C-ABI contract (include/caduceus_b200.h): out += silu(z) * sum_n C exp2(A2 cumdt) h0, per
        job (whole shard, carry a.h0) or per (job, logical segment) with the carries of cad_seg_carry.  bf16 I/O."""
        a = ref._obj
        self.calls.append(f"fixup nseg {a.nseg} first {a.seg_first}")
        assert a.io_dtype == _lib.CAD_BF16
        E, L, nj = a.E, a.L, a.njobs
        bf = lambda ptr, shape: (_arr(ptr, shape, np.uint16).astype(np.uint32) << 16).view(np.float32)   # noqa: E731
        xz, delta = bf(a.xz, (a.nseq, 2 * E, a.ldxz)), bf(a.delta, (nj, E, a.ldd))
        out_raw = _arr(a.out, (nj, E, a.ldo), np.uint16)
        out = (out_raw.astype(np.uint32) << 16).view(np.float32).astype(np.float64)
        bc = _arr(a.bc, (nj, 2 * N, a.ldbc))
        seq, ps, rv = (_arr(p, (nj,), np.int32) for p in (a.seq_of_job, a.pset_of_job, a.rev_of_job))
        P = int(ps.max()) + 1
        dt_b, A2 = _arr(a.dt_b, (P, E)), _arr(a.A2, (P, E, N))
        nseg = a.nseg if a.nseg > 1 else 1
        nch = (L + 255) // 256
        per = (nch + nseg - 1) // nseg
        for j in range(nj):
            for sl in range(0 if (nseg == 1 or a.seg_first) else 1, nseg):
                k = nseg - 1 - sl if rv[j] else sl
                lo, hi = min(k * per * 256, L), min((k + 1) * per * 256, L)
                if hi <= lo:
                    continue
                h0 = (_arr(a.seg_carry, (nj, nseg, E, N))[j, sl] if nseg > 1 else _arr(a.h0, (nj, E, N))[j]).astype(np.float64)
                idx = np.arange(hi - 1, lo - 1, -1) if rv[j] else np.arange(lo, hi)
                dt = _softplus(delta[j][:, idx].astype(np.float64) + dt_b[ps[j]].astype(np.float64)[:, None])
                cum = np.cumsum(dt, axis=1)
                Cm = bc[j, N:][:, idx].astype(np.float64)
                z = xz[seq[j], E:][:, idx].astype(np.float64)
                term = np.einsum("nt,ent,en->et", Cm, np.exp2(A2[ps[j]].astype(np.float64)[:, :, None] * cum[:, None, :]), h0)
                out[j][:, idx] += term * _silu(z)
        out_raw[:] = torch.from_numpy(out.astype(np.float32)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
        return 0


@pytest.fixture()
def host(emu, monkeypatch):   # noqa: F811
    lib = _EmuLib(emu)
    monkeypatch.setattr(_lib, "load", lambda: lib)
    monkeypatch.setattr(CF, "_stream", lambda: None)
    return lib


def _inputs(L, E, spec, seed):
    xz, delta, bc, conv_w4, conv_b, dt_b, A2, Dk, tabs, ld, ldbc = _problem(L, E, spec, torch.bfloat16, seed)
    return xz, delta, bc, (conv_w4, conv_b, dt_b, A2, Dk), tuple(tabs)


def _ref(xz, delta, bc, packed, spec, L, **kw):
    f = lambda t: t.float().numpy()   # noqa: E731
    return boundary_ref(f(xz), f(delta), f(bc), *(f(t) for t in packed), [s for s, _, _ in spec], [q for _, q, _ in spec],
                        [r for _, _, r in spec], L, **kw)


def _close(got, ref, part=None, atol=2e-4):
    # two roundings to bf16 where a carry term was added: up to one ulp of the larger of (result, partial result)
    scale = np.abs(ref) if part is None else np.maximum(np.abs(ref), np.abs(part))
    err, bound = np.abs(got - ref), atol + 2.1 * 2.0 ** -8 * scale
    assert (err <= bound).all(), (err.max(), (err - bound).max())


@pytest.mark.parametrize("nseg", [None, 1, 3])
def test_host_flow_plain_inference(host, nseg):
    L, E, spec = 4500, 40, [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)]
    xz, delta, bc, packed, jobs = _inputs(L, E, spec, 11)
    out, hl, ds, ctx = CF.scan_fwd(xz, delta, bc, packed, jobs, L, variant=20, nseg=nseg)
    assert hl is None and ds is None and ctx is None
    want = CF.default_nseg(4, E, L, 2) if nseg is None else nseg          # 4 jobs, E = 40 -> 2 warps per CTA
    assert host.calls[0] == "transpose" and host.calls[1] == f"scan v20 nseg {want} W 2"
    assert host.calls[2:] == (["carry", f"fixup nseg {want} first 0"] if want > 1 else [])
    # (the zero-carry partial result is internal to the flow here: where it cancelled against the carry term its ulp
    #  exceeds the result's, hence the absolute slack when there is more than one segment)
    _close(out[..., :L].float().numpy(), _ref(xz, delta, bc, packed, spec, L), atol=2e-4 if want == 1 else 8e-3)


def test_host_flow_token_major_copy_supplied_by_conv_xproj(host):
    L, E, spec = 1300, 32, [(0, 0, 0), (0, 1, 1)]
    xz, delta, bc, packed, jobs = _inputs(L, E, spec, 12)
    bcT = torch.zeros(2, 1536, 2 * N)
    bcT[:, :L] = bc[..., :L].transpose(1, 2)
    out = CF.scan_fwd(xz, delta, bc, packed, jobs, L, variant=20, nseg=2, bcT=bcT.contiguous())[0]
    assert "transpose" not in host.calls and host.calls[0] == "scan v20 nseg 2 W 1"
    _close(out[..., :L].float().numpy(), _ref(xz, delta, bc, packed, spec, L), atol=8e-3)


@pytest.mark.parametrize("nseg", [1, 4])
def test_host_flow_as_a_sequence_shard(host, nseg):
    """modules.py's sharded branch: zero-carry scan with halo -> (hlast, dtsum) for the all_gather -> scan_fixup(h0, seg_ctx)."""
    L, E, spec = 2300, 40, [(0, 0, 1), (0, 1, 0)]
    xz, delta, bc, packed, jobs = _inputs(L, E, spec, 13)
    g = torch.Generator().manual_seed(1)
    halo, h0 = torch.randn(2, E, 3, generator=g).to(torch.bfloat16), torch.randn(2, E, N, generator=g)
    out, hl, ds, ctx = CF.scan_fwd(xz, delta, bc, packed, jobs, L, halo=halo, want_state=True, variant=20, nseg=nseg)
    zero = _ref(xz, delta, bc, packed, spec, L, halo=halo.float().numpy(), full=True)
    assert np.allclose(hl.numpy(), zero[1], rtol=2e-4, atol=2e-4 * max(1.0, np.abs(zero[1]).max()))
    assert np.allclose(ds.numpy(), zero[2], rtol=2e-4, atol=1e-4)
    assert isinstance(ctx, dict) and ctx["nseg"] == nseg and "fixup" not in " ".join(host.calls)
    part = out[..., :L].float().numpy().copy()
    CF.scan_fixup(xz, delta, bc, out, packed, jobs, L, h0, seg_ctx=ctx)
    assert host.calls[-1] == (f"fixup nseg {nseg} first 1" if nseg > 1 else "fixup nseg 0 first 0")
    _close(out[..., :L].float().numpy(), _ref(xz, delta, bc, packed, spec, L, halo=halo.float().numpy(), h0=h0.numpy()), part)
