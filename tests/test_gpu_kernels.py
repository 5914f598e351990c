"""GPU parity: each CUDA primitive of libsetok_b200 (called through the C ABI) against torch fp32 on the
same operands.  Tolerances: fp32 outputs 1e-4 of the output scale (identical bf16 operands, fp32
accumulation: only summation order differs); bf16 outputs one bf16 ulp (2^-8 relative) on top."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

from setok_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")


def _close(got, ref, rel, what=""):
    got, ref = got.float(), ref.float()
    scale = ref.abs().max().clamp_min(1e-6)
    err = (got - ref).abs().max() / scale
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    assert err <= rel, f"{what}: normalised max error {err:.3e} > {rel:.1e}"


def _gemm_ref(a, w, bias, act, residual):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == ops.ACT_QUICK_GELU:
        y = y * torch.sigmoid(1.702 * y)
    elif act == ops.ACT_GELU_ERF:
        y = torch.nn.functional.gelu(y)
    if residual is not None:
        y = y + residual.float()
    return y


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (257, 1024, 1024), (1000, 3072, 1024), (514, 1024, 4096), (77, 64, 64),
                                   (300, 48, 96), (1, 8, 8), (2056, 4096, 640), (129, 264, 72)])
def test_gemm_shapes(M, N, K):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, generator=g)).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    out = ops.gemm(a, w, bias, out_dtype=torch.float32)
    _close(out, _gemm_ref(a, w, bias, 0, None), 1e-4, f"gemm f32 {M}x{N}x{K}")
    out = ops.gemm(a, w, bias, out_dtype=torch.bfloat16)
    _close(out, _gemm_ref(a, w, bias, 0, None), 5e-3, f"gemm bf16 {M}x{N}x{K}")


@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_QUICK_GELU, ops.ACT_GELU_ERF])
@pytest.mark.parametrize("res", [None, torch.bfloat16, torch.float32])
def test_gemm_epilogues(act, res):
    M, N, K = 391, 520, 256
    g = torch.Generator().manual_seed(5)
    a = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    r = None if res is None else torch.randn(M, N, generator=g).to(DEV, res)
    ref = _gemm_ref(a, w, bias, act, r)
    _close(ops.gemm(a, w, bias, act=act, residual=r, out_dtype=torch.float32), ref, 1e-4, "epilogue f32")
    _close(ops.gemm(a, w, None, act=act, residual=r, out_dtype=torch.float32), _gemm_ref(a, w, None, act, r), 1e-4, "no bias")
    if r is not None:     # in place (the residual stream is updated in place on the product path)
        out = r.clone()
        ops.gemm(a, w, bias, act=act, residual=out, out=out)
        _close(out, ref, 5e-3 if res == torch.bfloat16 else 1e-4, "in-place residual")


@pytest.mark.parametrize("act,res,odt", [(ops.ACT_NONE, None, torch.bfloat16), (ops.ACT_QUICK_GELU, None, torch.bfloat16),
                                         (ops.ACT_GELU_ERF, None, torch.bfloat16), (ops.ACT_NONE, torch.bfloat16, torch.bfloat16),
                                         (ops.ACT_NONE, torch.float32, torch.float32), (ops.ACT_NONE, None, torch.float32),
                                         (ops.ACT_NONE, torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("M,N,K", [(391, 512, 256), (2056, 1024, 2048), (77, 96, 64), (1, 32, 8), (700, 3072, 1024)])
def test_gemm_row_owner_epilogue_equals_transposing_epilogue(act, res, odt, M, N, K):
    """The two epilogues of the tcgen05 GEMM (shared-memory transpose + coalesced stores / row-owner 256-bit stores) perform
    the same float operations in the same order: their outputs must be bit-identical, ragged rows, in-place residual and the
    device-side row count included, and rows past the live count stay untouched."""
    import ctypes
    from setok_b200 import _lib
    lib = _lib.load()
    lib.setok_debug_set_gemm_epi_direct.argtypes = [ctypes.c_int]
    lib.setok_debug_set_gemm_epi_direct.restype = None
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * K ** -0.5).to(DEV, torch.bfloat16)
    bias = torch.randn(N, generator=g).to(DEV)
    r = None if res is None else torch.randn(M, N, generator=g).to(DEV, res)
    live = max(1, M - 37)
    m_dev = torch.tensor([live], dtype=torch.int32, device=DEV)
    outs = []
    try:
        for mode in (0, 1):
            lib.setok_debug_set_gemm_epi_direct(mode)
            full = ops.gemm(a, w, bias, act=act, residual=r, out_dtype=odt)
            part = torch.full((M, N), -7.0, device=DEV, dtype=odt)
            ops.gemm(a, w, bias, act=act, residual=r, out=part, m_dev=m_dev)
            inplace = None
            if r is not None and res == odt:
                inplace = r.clone()
                ops.gemm(a, w, bias, act=act, residual=inplace, out=inplace)
            outs.append((full, part, inplace))
    finally:
        lib.setok_debug_set_gemm_epi_direct(-1)
    (f0, p0, i0), (f1, p1, i1) = outs
    assert torch.equal(f0, f1) and torch.equal(p0, p1)
    assert torch.equal(p1[:live], f1[:live]) and bool((p1[live:] == -7.0).all())
    if i0 is not None:
        assert torch.equal(i0, i1) and torch.equal(i1, f1)
    _close(f1, _gemm_ref(a, w, bias, act, r), 1e-4 if odt == torch.float32 else 5e-3, "row-owner epilogue")


def test_gemm_device_row_count_and_untouched_tail():
    M, N, K = 1000, 256, 128
    g = torch.Generator().manual_seed(9)
    a = torch.randn(M, K, generator=g).to(DEV, torch.bfloat16)
    w = torch.randn(N, K, generator=g).to(DEV, torch.bfloat16)
    for live in (0, 1, 127, 128, 129, 777, 1000, 5000):
        out = torch.full((M, N), -7.0, device=DEV)
        m_dev = torch.tensor([live], dtype=torch.int32, device=DEV)
        ops.gemm(a, w, None, out=out, m_dev=m_dev)
        n = min(live, M)
        _close(out[:n], _gemm_ref(a[:n], w, None, 0, None), 1e-4, f"m_dev={live}") if n else None
        assert (out[n:] == -7.0).all(), f"rows past the live count were written (m_dev={live})"


def test_gemm_linearity_full_size():
    """Size-independent property at a BASELINE-size GEMM (65792 x 1024 x 1024): f(a1 + a2) == f(a1) + f(a2)."""
    M, N, K = 65792, 1024, 1024
    g = torch.Generator().manual_seed(1)
    a1 = torch.randint(-4, 5, (M, K), generator=g).to(DEV, torch.bfloat16)     # small integers: exact in bf16 and fp32
    a2 = torch.randint(-4, 5, (M, K), generator=g).to(DEV, torch.bfloat16)
    w = torch.randint(-2, 3, (N, K), generator=g).to(DEV, torch.bfloat16)
    y1 = ops.gemm(a1, w, out_dtype=torch.float32)
    y2 = ops.gemm(a2, w, out_dtype=torch.float32)
    y12 = ops.gemm((a1.float() + a2.float()).to(torch.bfloat16), w, out_dtype=torch.float32)
    assert torch.equal(y12, y1 + y2)
    rows = torch.randint(0, M, (64,), generator=g).to(DEV)
    assert torch.equal(y1[rows], a1[rows].float() @ w.float().t())


@pytest.mark.parametrize("rows,C", [(5, 64), (300, 768), (1000, 1024), (37, 4096), (16, 48)])
@pytest.mark.parametrize("dt_in,dt_out", [(torch.float32, torch.bfloat16), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32)])
def test_layernorm(rows, C, dt_in, dt_out):
    g = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=g) * 3 + 1).to(DEV, dt_in)
    gamma = (1 + 0.1 * torch.randn(C, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(C, generator=g)).to(DEV)
    ref = torch.nn.functional.layer_norm(x.float(), (C,), gamma, beta, 1e-5)
    out = ops.layernorm(x, gamma, beta, 1e-5, out_dtype=dt_out)
    _close(out, ref, 5e-3 if dt_out == torch.bfloat16 else 1e-5, "layernorm")
    perm = torch.randperm(rows, generator=g).to(DEV, torch.int32)
    out = ops.layernorm(x, gamma, beta, 1e-5, out_dtype=dt_out, gather=perm)
    _close(out, ref[perm.long()], 5e-3 if dt_out == torch.bfloat16 else 1e-5, "layernorm gather")


def _attn_ref(qkv, heads, scale, seg_off):
    rows, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    q, k, v = qkv.float().split(C, dim=1)
    out = torch.zeros(rows, C, device=qkv.device)
    so = seg_off.tolist()
    for s in range(len(so) - 1):
        a, b = so[s], so[s + 1]
        if a == b:
            continue
        for h in range(heads):
            sl = slice(h * hd, (h + 1) * hd)
            att = torch.softmax((q[a:b, sl] @ k[a:b, sl].t()) * scale, dim=-1)
            out[a:b, sl] = att @ v[a:b, sl]
    return out


@pytest.mark.parametrize("C,heads", [(64, 2), (128, 2), (768, 2), (1024, 2), (64, 4)])
def test_attention_ragged_segments(C, heads):
    g = torch.Generator().manual_seed(C + heads)
    lens = [1, 7, 130, 2, 33, 1, 64, 5]
    so = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32)
    rows = int(so[-1])
    row_seg = torch.repeat_interleave(torch.arange(len(lens)), torch.tensor(lens)).to(torch.int32)
    qkv = torch.randn(rows, 3 * C, generator=g).to(DEV, torch.bfloat16)
    scale = (C // heads) ** -0.5
    out = ops.attention(qkv, heads, scale, seg_off=so.to(DEV), row_seg=row_seg.to(DEV))
    _close(out, _attn_ref(qkv, heads, scale, so), 6e-3, "ragged attention")
    # device-side live row count: rows past it untouched by construction of the kernel loop
    m_dev = torch.tensor([138], dtype=torch.int32, device=DEV)
    out2 = ops.attention(qkv, heads, scale, seg_off=so.to(DEV), row_seg=row_seg.to(DEV), m_dev=m_dev)
    assert torch.equal(out2[:138], out[:138])


@pytest.mark.parametrize("T,B,heads", [(17, 3, 2), (64, 2, 4), (82, 2, 2), (197, 2, 12), (257, 3, 16), (577, 1, 16),
                                       # T <= 257 runs the whole-row kernel (attention_fullrow.cu): one / two query tiles, key counts that
                                       # are not multiples of 16 or 64, the 257th key + query row, and more (image, head) pairs than SMs so
                                       # that every persistent CTA walks several pairs through its 2-stage ring (barrier parities wrap)
                                       (1, 4, 2), (16, 5, 2), (128, 10, 16), (129, 10, 16), (256, 10, 16), (197, 30, 12), (257, 20, 16), (257, 40, 16),
                                       (1025, 1, 2)])
def test_attention_vit_hd64(T, B, heads):
    C = heads * 64
    g = torch.Generator().manual_seed(T)
    qkv = torch.randn(B * T, 3 * C, generator=g).to(DEV, torch.bfloat16)
    so = torch.arange(0, B * T + 1, T, dtype=torch.int32)
    out = ops.attention(qkv, heads, 0.125, uniform_T=T)
    ref = _attn_ref(qkv, heads, 0.125, so)
    _close(out, ref, 6e-3, "vit attention")
    # both kernels on every shape they accept (the dispatcher picks the whole-row kernel only for 225 <= T <= 257)
    import ctypes
    from setok_b200 import _lib
    lib = _lib.load()
    lib.setok_debug_set_attention_fullrow.argtypes = [ctypes.c_int]
    try:
        for mode in (2, 0):
            lib.setok_debug_set_attention_fullrow(mode)
            _close(ops.attention(qkv, heads, 0.125, uniform_T=T), ref, 6e-3, f"vit attention (fullrow mode {mode})")
    finally:
        lib.setok_debug_set_attention_fullrow(1)


@pytest.mark.parametrize("Bt,M,N,K,mn", [(3, 256, 256, 512, False), (5, 196, 196, 384, False), (4, 256, 512, 256, True), (3, 196, 384, 196, True), (2, 576, 512, 576, True),
                                         (2, 100, 64, 72, True), (1, 300, 264, 128, False)])
def test_gemm_batched_and_mn_major(Bt, M, N, K, mn):
    """Batched products with strided operands (views into a packed qkv buffer) and the [K, N] (MN-major) B form."""
    g = torch.Generator().manual_seed(Bt * 1000 + M + N + K)
    pad = 24 + (-K % 8)                                          # leading dimensions stay multiples of 8 elements
    a_full = torch.randn(Bt, M, K + pad, generator=g).to(DEV, torch.bfloat16)
    a = a_full[:, :, 8:8 + K]                                     # row stride K+pad, 16-byte aligned start
    if mn:
        w_full = (torch.randn(Bt, K, N + 24 + (-N % 8), generator=g) * K ** -0.5).to(DEV, torch.bfloat16)
        w = w_full[:, :, 16:16 + N]
        ref = torch.bmm(a.float(), w.float())
    else:
        w_full = (torch.randn(Bt, N, K + pad, generator=g) * K ** -0.5).to(DEV, torch.bfloat16)
        w = w_full[:, :, 8:8 + K]
        ref = torch.bmm(a.float(), w.float().transpose(1, 2))
    out = ops.gemm_batched(a, w, w_mn_major=mn, out_dtype=torch.float32)
    _close(out, ref, 1e-4, f"batched gemm mn={mn}")
    out = ops.gemm_batched(a, w, w_mn_major=mn, out_dtype=torch.bfloat16)
    _close(out, ref, 5e-3, f"batched gemm bf16 mn={mn}")


# ---------------------------------------------------------------------------------------------------
# LayerNorm folded into the GEMMs around it (setok_gemm_bf16_ln, setok_ln_fold_init)
# ---------------------------------------------------------------------------------------------------
def _fold_pack(w0, b0, gamma, beta):
    """What the tower packs under SETOK_VIT_LN_FOLD (include/setok_b200.h): W' = bf16(gamma (.) W0), s = row sums of W', t = W0 beta + b0."""
    wg = (w0 * gamma[None, :]).to(torch.bfloat16).contiguous()
    return wg, wg.float().sum(1).contiguous(), ((w0 * beta[None, :]).sum(1) + b0).contiguous()


@pytest.mark.parametrize("M,C,N2,K2", [(300, 1024, 3072, 1024), (65, 128, 512, 128), (700, 768, 3072, 3072), (160, 96, 256, 64)])
@pytest.mark.parametrize("act", [ops.ACT_NONE, ops.ACT_QUICK_GELU])
def test_gemm_layernorm_fold_chain(M, C, N2, K2, act):
    """x0 -> init -> consumer (LN(x0) W1^T + b1) ; producer (x1 = x0 + a W2^T + b2, xhat / records of x1) -> consumer on x1, against
    torch fp32 LayerNorm + matmul on the same bf16-rounded operands.  The stream has a large row mean and per-row scale so that
    the lagged statistics (c, r) of the producer differ visibly from the final ones; C = 96 leaves the last 128-column slot
    partial, K2 = 3072 takes the row-owner producer epilogue (K >= 2048), the others the transposing one."""
    g = torch.Generator(device=DEV).manual_seed(M + C + act)
    eps = 1e-5
    x0 = (torch.randn(M, C, device=DEV, generator=g) * (0.5 + torch.rand(M, 1, device=DEV, generator=g) * 3) + torch.randn(M, 1, device=DEV, generator=g) * 2).contiguous()
    gamma = 1.0 + 0.3 * torch.randn(C, device=DEV, generator=g)
    beta = 0.3 * torch.randn(C, device=DEV, generator=g)
    w1 = torch.randn(N2, C, device=DEV, generator=g) * C ** -0.5
    b1 = 0.1 * torch.randn(N2, device=DEV, generator=g)
    wg, s1, t1 = _fold_pack(w1, b1, gamma, beta)

    def consumer_ref(x):
        ln = torch.nn.functional.layer_norm(x, (C,), gamma, beta, eps)
        y = ln @ w1.t() + b1
        return y * torch.sigmoid(1.702 * y) if act == ops.ACT_QUICK_GELU else y

    xhat, rec = ops.ln_fold_init(x0, eps)
    y0 = ops.gemm_ln(xhat, wg, t1, rec, ln_C=C, eps=eps, act=act, ln_s=s1)
    _close(y0, consumer_ref(x0), 1.2e-2, "consumer after init")
    # producer: x1 = x0 + a W2^T + b2 (in place), with a shift of the row mean and a change of scale
    a = torch.randn(M, K2, device=DEV, generator=g).to(torch.bfloat16)
    w2 = (torch.randn(C, K2, device=DEV, generator=g) * (2.0 * K2 ** -0.5)).to(torch.bfloat16)
    b2 = 0.5 + 0.5 * torch.randn(C, device=DEV, generator=g)
    x1_ref = x0 + a.float() @ w2.float().t() + b2
    x1 = x0.clone()
    rec1 = ops.ln_records(M, C, DEV)
    xhat1 = torch.empty(M, C, dtype=torch.bfloat16, device=DEV)
    ops.gemm_ln(a, w2, b2, rec, ln_C=C, eps=eps, residual=x1, rec_out=rec1, xhat=xhat1, out=x1)
    _close(x1, x1_ref, 1e-4, "producer stream")
    # the record reconstructs the row statistics of x1 exactly (fp32) ...
    c, r = rec1[:, 0], rec1[:, 1]
    S1, S2 = rec1[:, 2::2].sum(1), rec1[:, 3::2].sum(1)
    mu = c + S1 / C
    var = S2 / C - (S1 / C) ** 2
    assert torch.allclose(mu, x1_ref.mean(1), rtol=1e-4, atol=1e-4)
    assert torch.allclose(var, x1_ref.var(1, unbiased=False), rtol=2e-3, atol=1e-5)
    # ... (c, r) are the statistics of the row BEFORE the update, and xhat is (x1 - c) r rounded to bf16
    assert torch.allclose(c, x0.mean(1), rtol=1e-4, atol=1e-4)
    assert torch.allclose(r, (x0.var(1, unbiased=False) + eps).rsqrt(), rtol=1e-3)
    _close(xhat1, (x1_ref - c[:, None]) * r[:, None], 2 ** -8, "xhat")
    y1 = ops.gemm_ln(xhat1, wg, t1, rec1, ln_C=C, eps=eps, act=act, ln_s=s1)
    _close(y1, consumer_ref(x1_ref), 1.2e-2, "consumer after producer")
    # a second producer on top (records ping-pong back), checks the chain c' = c + m
    x2_ref = x1_ref + a.float() @ w2.float().t() + b2
    ops.gemm_ln(a, w2, b2, rec1, ln_C=C, eps=eps, residual=x1, rec_out=rec, xhat=xhat, out=x1)
    _close(x1, x2_ref, 1e-4, "second producer stream")
    assert torch.allclose(rec[:, 0], x1_ref.mean(1), rtol=1e-4, atol=1e-4)
    y2 = ops.gemm_ln(xhat, wg, t1, rec, ln_C=C, eps=eps, act=act, ln_s=s1)
    _close(y2, consumer_ref(x2_ref), 1.2e-2, "consumer after the second producer")


def test_gemm_layernorm_fold_rejects_bad_layouts():
    from setok_b200 import SetokError
    x = torch.randn(64, 100, device=DEV)                    # C % 32 != 0
    xhat, rec = ops.ln_fold_init(x)
    w = torch.randn(64, 104, device=DEV).to(torch.bfloat16)[:, :100]
    with pytest.raises(SetokError):
        ops.gemm_ln(xhat, w, torch.zeros(64, device=DEV), rec, ln_C=100, ln_s=None)      # neither side selected
