"""tcgen05 linear-layer kernel (ug_gemm_bf16) against an fp32 torch reference of the same op.

Tolerance (stated, bf16 outputs of an fp32-accumulated product): relative L2 error <= 4e-3 and max-abs error
<= 2^-6 * max|ref| against the fp32 reference -- one bf16 rounding (2^-9 relative) plus accumulation-order noise;
cuBLAS' own bf16 output is held to the same bound in the same test."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from ml_unigen_b200 import ops
    return ops


def _check(got, ref32):
    d = got.float() - ref32
    rel = (d.norm() / ref32.norm()).item()
    mx = d.abs().max().item()
    assert rel <= 4e-3, rel
    assert mx <= ref32.abs().max().item() * 2 ** -6, mx


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 128), (300, 512, 1536), (1000, 2048, 1536),
                                   (2082, 1536, 8960), (37, 256, 192)])
@pytest.mark.parametrize("with_bias", [False, True])
def test_gemm_plain(M, N, K, with_bias):
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M * 7 + N + K)
    x = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    b = torch.randn(N, device=dev, generator=g).bfloat16() if with_bias else None
    got = ops.gemm(x, w, b)
    ref32 = F.linear(x.float(), w.float(), None if b is None else b.float())
    _check(got, ref32)
    _check(F.linear(x, w, b), ref32)


def test_gemm_plain_strided_rows():
    """a and out as column slices of wider tensors (row strides != widths), as the backbone passes them."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    big = torch.randn(515, 2048, device=dev, generator=g).bfloat16()
    x = big[:, 256:256 + 1536]
    w = (torch.randn(512, 1536, device=dev, generator=g) * 0.03).bfloat16()
    out_big = torch.zeros(515, 1024, dtype=torch.bfloat16, device=dev)
    ops.gemm(x, w, out=out_big[:, 512:])
    _check(out_big[:, 512:], F.linear(x.float(), w.float()))
    assert out_big[:, :512].abs().max().item() == 0.0


@pytest.mark.parametrize("M,I,K", [(128, 128, 64), (300, 256, 1536), (2082, 8960, 1536), (16 * 130, 1280, 1536),
                                   (16 * 258, 8960, 1536), (5000, 1024, 256), (16 * 1026, 1152, 128)])
def test_gemm_swiglu(M, I, K):
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M + I + K)
    x = torch.randn(M, K, device=dev, generator=g).bfloat16()
    wg = (torch.randn(I, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    wu = (torch.randn(I, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    got = ops.gemm(x, ops.pack_gate_up(wg, wu), mode=ops.GEMM_SWIGLU)
    assert got.shape == (M, I)
    gate = F.linear(x.float(), wg.float())
    up = F.linear(x.float(), wu.float())
    ref32 = F.silu(gate) * up
    d = got.float() - ref32
    # two bf16 roundings of the factors plus one of the product
    assert (d.norm() / ref32.norm()).item() <= 8e-3
    assert d.abs().max().item() <= ref32.abs().max().item() * 2 ** -5
    # and against the unfused path (cuBLAS GEMM + ug_swiglu): same arithmetic after the GEMM, so almost every element equal
    unf = ops.swiglu(F.linear(x, torch.cat([wg, wu], 0)))
    assert ((got.float() - unf.float()).norm() / unf.float().norm()).item() <= 4e-3


@pytest.mark.parametrize("M,N", [(16 * 1026, 1536), (16 * 1026, 2048), (9000, 768), (700, 256)])
def test_gemm_tail_split_matches_cublas(M, N):
    """Shapes whose last round is cut into 256 x 128 half tiles (N = 128 MMAs): every output element still equals the
    fp32 reference within the bf16 bound, including the rows of the ragged last m block."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M + N)
    K = 512
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    got = ops.gemm(a, w)
    ref = F.linear(a, w)
    assert ((got.float() - ref.float()).abs().max().item()) <= 2 ** -6 * ref.float().abs().max().item()
    _check(got, F.linear(a.float(), w.float()))


@pytest.mark.parametrize("B,L,Hq,Hkv,K,with_cache", [(2, 111, 12, 2, 1536, True), (3, 70, 2, 1, 256, True), (1, 300, 12, 2, 1536, False),
                                                        (16, 130, 12, 2, 1536, True), (2, 257, 6, 1, 128, True)])
def test_qkv_rope_fused_matches_unfused(B, L, Hq, Hkv, K, with_cache):
    """ug_qkv_rope_bf16 against the unfused sequence it replaces (plain GEMM + ug_rope_qk) and against the fp32 formula
    of transformers' apply_rotary_pos_emb.  The rotation may contract its multiply-adds differently from the row kernel,
    so equality is demanded within one bf16 ulp and on >= 99.9 % of the elements exactly."""
    from oracle import backbone_oracle as BO
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(B * 1000 + L)
    D, M = 128, B * L
    N = (Hq + 2 * Hkv) * D
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device=dev, generator=g).bfloat16()
    Lc, row0 = L + 37, 37
    cos, sin = BO.rope_cos_sin(Lc, D, 1e6, dev)
    cos_w, sin_w = cos[row0:, : D // 2].contiguous(), sin[row0:, : D // 2].contiguous()      # positions row0 .. row0+L-1
    cache_f = torch.zeros(B, Lc, 2, Hkv, D, dtype=torch.bfloat16, device=dev) if with_cache else None
    cache_u = torch.zeros_like(cache_f) if with_cache else None
    fused = ops.qkv_rope(a, w, bias, cos_w, sin_w, L, Hq, Hkv, kv_cache=cache_f, cache_row_offset=row0)
    unf = ops.gemm(a, w, bias)
    ops.rope_qk_(unf, cos_w, sin_w, L, Hq, Hkv, D, kv_cache=cache_u, cache_row_offset=row0)
    d = (fused.float() - unf.float()).abs()
    assert (d <= unf.float().abs() * 2 ** -7 + 1e-6).all()
    assert (fused == unf).float().mean().item() >= 0.999
    assert torch.equal(fused[:, (Hq + Hkv) * D:], unf[:, (Hq + Hkv) * D:])          # v heads: untouched by the rotation
    if with_cache:
        assert torch.equal(cache_f[:, row0:].reshape(M, 2 * Hkv * D), fused[:, Hq * D:])
        assert int(cache_f[:, :row0].abs().sum()) == 0
    # fp32 formula
    x = F.linear(a.float(), w.float(), bias.float())
    q = x[:, : Hq * D].view(B, L, Hq, D)
    ref_q = BO.apply_rope(q, cos[row0:], sin[row0:])
    got_q = fused[:, : Hq * D].view(B, L, Hq, D).float()
    assert ((got_q - ref_q).norm() / ref_q.norm()).item() <= 6e-3


def test_gemm_rejects_bad_shapes():
    ops = _ops()
    from ml_unigen_b200._lib import UnigenB200Error
    dev = torch.device("cuda:0")
    x = torch.zeros(64, 100, dtype=torch.bfloat16, device=dev)
    w = torch.zeros(256, 100, dtype=torch.bfloat16, device=dev)
    with pytest.raises((UnigenB200Error, ValueError)):
        ops.gemm(x, w)
    x = torch.zeros(64, 128, dtype=torch.bfloat16, device=dev)
    w = torch.zeros(200, 128, dtype=torch.bfloat16, device=dev)
    with pytest.raises((UnigenB200Error, ValueError)):
        ops.gemm(x, w)


def test_gemm_shape_sweep_against_cublas():
    """Seeded sweep over ragged M, every tail-round regime (no tail, whole tiles, half tiles), small and large K, all three
    epilogues, against cuBLAS within one bf16 rounding (the accumulation order along K is the same as that of the
    cuBLAS kernels seen so far, so most outputs are in fact bit-identical: printed)."""
    import random
    ops = _ops()
    dev = torch.device("cuda:0")
    rng = random.Random(7)
    g = torch.Generator(device=dev).manual_seed(7)
    n_equal = 0
    for case in range(36):
        M = rng.choice([1, 7, 128, 129, 255, 256, 257, 1000, 4128, 5000, 9473, 19001])
        N = 256 * rng.choice([1, 2, 3, 6, 8, 13, 37])
        K = 64 * rng.choice([1, 2, 3, 8, 24, 47])
        a = torch.randn(M, K, device=dev, generator=g).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
        b = torch.randn(N, device=dev, generator=g).bfloat16() if case % 2 else None
        got = ops.gemm(a, w, b)
        ref = F.linear(a, w, b)
        d = (got.float() - ref.float()).abs().max().item()
        assert d <= 2 ** -6 * max(ref.float().abs().max().item(), 1e-3), (M, N, K, d)
        n_equal += int(torch.equal(got, ref))
        if N % 512 == 0 and case % 3 == 0:
            I = N // 2
            if I % 128 == 0:
                act = ops.gemm(a, ops.pack_gate_up(w[:I], w[I:]), mode=ops.GEMM_SWIGLU)
                unf = ops.swiglu(F.linear(a, w))
                assert ((act.float() - unf.float()).abs().max().item()) <= 2 ** -6 * max(unf.float().abs().max().item(), 1e-3), (M, N, K)
    print(f"{n_equal} of 36 plain outputs bit-identical to cuBLAS")      # informational: depends on cuBLAS' kernel choice


# ------------------------------------------------------------------------------------------------------------
# round 2: o_proj / down_proj + residual add + RMSNorm statistics in one launch; RMSNorm folded into the consumer GEMMs
@pytest.mark.parametrize("M,N,K", [(300, 512, 1536), (2082, 1536, 8960), (16 * 1026, 1536, 1536), (37, 256, 192), (129, 1536, 64),
                                   (700, 256, 512), (9000, 768, 512)])
def test_gemm_residual_epilogue(M, N, K):
    """ug_gemm_residual_bf16: x += bf16(a w^T), xb = bf16(x), ss = per-128-column sums of x^2.  The accumulator is the plain
    mode's, so x must equal x0 + float(ops.gemm(a, w)) BIT FOR BIT; xb is its bf16 rounding exactly; ss is an fp32 sum
    (different order from torch's: 1e-5 relative).  Covers the split-tail rounds and ragged last m blocks."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    x0 = torch.randn(M, N, device=dev, generator=g) * 3
    x = x0.clone()
    guard = torch.full((64,), 7.0, device=dev)                  # allocated right behind: a write past row M would show
    xb, ss = ops.gemm_residual_(a, w, x)
    want = x0 + ops.gemm(a, w).float()
    assert torch.equal(x, want)
    assert torch.equal(xb, want.bfloat16())
    ref_ss = (want.double() ** 2).view(M, N // 128, 128).sum(-1)
    assert ((ss.double() - ref_ss).abs() <= 1e-5 * ref_ss + 1e-12).all()
    assert torch.equal(guard, torch.full((64,), 7.0, device=dev))
    # a second application accumulates on top (in-place semantics)
    xb2, ss2 = ops.gemm_residual_(a, w, x, xb, ss)
    assert xb2.data_ptr() == xb.data_ptr() and torch.equal(x, want + ops.gemm(a, w).float())


@pytest.mark.parametrize("M,K", [(300, 1536), (16 * 130, 1536), (777, 256)])
def test_rmsnorm_folding_matches_unfused_norm_then_gemm(M, K):
    """(g x / rms) W^T == (1/rms) (bf16(x) (W diag g)^T): the three consumer epilogues with `row_ss` against the unfused sequence
    ug_add_rmsnorm -> GEMM and against the fp32 formula.  The rounding points move (x is rounded to bf16 before the scaling
    instead of after, the weights carry the gain), so the two bf16 paths differ by rounding noise; both are held to the same
    bound against fp32: relative L2 <= 6e-3 (plain / q/k/v), 1e-2 (SwiGLU)."""
    from oracle import backbone_oracle as BO
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(M + K)
    eps = 1e-6
    x = torch.randn(M, K, device=dev, generator=g) * (0.5 + 4 * torch.rand(M, 1, device=dev, generator=g))   # rows of different scale
    gain = (1 + 0.2 * torch.randn(K, device=dev, generator=g)).bfloat16()
    xb, ss = ops.norm_operands(x)
    h_ref32 = gain.float() * (x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps))
    h_unf = ops.add_rmsnorm(x.clone(), gain, eps)
    fold = lambda w: (w.float() * gain.float()[None, :]).bfloat16()
    rel = lambda got, ref: ((got.float() - ref).norm() / ref.norm()).item()
    # plain (+bias): the codebook head
    N = 512
    w = (torch.randn(N, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    b = torch.randn(N, device=dev, generator=g).bfloat16()
    ref = F.linear(h_ref32, w.float(), b.float())
    got = ops.gemm(xb, fold(w), b, row_ss=ss, rms_eps=eps)
    unf = ops.gemm(h_unf, w, b)
    print(f"plain: folded {rel(got, ref):.2e}  unfused {rel(unf, ref):.2e}")
    assert rel(got, ref) <= 6e-3 and rel(unf, ref) <= 6e-3
    # SwiGLU
    I = 384
    wg = (torch.randn(I, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    wu = (torch.randn(I, K, device=dev, generator=g) * K ** -0.5).bfloat16()
    ref = F.silu(F.linear(h_ref32, wg.float())) * F.linear(h_ref32, wu.float())
    got = ops.gemm(xb, ops.pack_gate_up(fold(wg), fold(wu)), mode=ops.GEMM_SWIGLU, row_ss=ss, rms_eps=eps)
    unf = ops.gemm(h_unf, ops.pack_gate_up(wg, wu), mode=ops.GEMM_SWIGLU)
    print(f"SwiGLU: folded {rel(got, ref):.2e}  unfused {rel(unf, ref):.2e}")
    assert rel(got, ref) <= 1e-2 and rel(unf, ref) <= 1e-2
    # q/k/v + RoPE
    if M % 3 == 0:
        B, L, Hq, Hkv, D = 3, M // 3, 2, 1, 128
        wq = (torch.randn((Hq + 2 * Hkv) * D, K, device=dev, generator=g) * K ** -0.5).bfloat16()
        bq = torch.randn((Hq + 2 * Hkv) * D, device=dev, generator=g).bfloat16()
        cos, sin = BO.rope_cos_sin(L, D, 1e6, dev)
        ch, sh = cos[:, : D // 2].contiguous(), sin[:, : D // 2].contiguous()
        got = ops.qkv_rope(xb, fold(wq), bq, ch, sh, L, Hq, Hkv, row_ss=ss, rms_eps=eps)
        unf = ops.qkv_rope(h_unf, wq, bq, ch, sh, L, Hq, Hkv)
        full = F.linear(h_ref32, wq.float(), bq.float())
        ref_q = BO.apply_rope(full[:, : Hq * D].view(B, L, Hq, D), cos, sin)
        gq = got[:, : Hq * D].view(B, L, Hq, D)
        uq = unf[:, : Hq * D].view(B, L, Hq, D)
        print(f"q/k/v: folded {rel(gq, ref_q):.2e}  unfused {rel(uq, ref_q):.2e}")
        assert rel(gq, ref_q) <= 6e-3 and rel(uq, ref_q) <= 6e-3
        assert rel(got[:, (Hq + Hkv) * D:], full[:, (Hq + Hkv) * D:]) <= 6e-3


def test_embed_rows_norm_operands():
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    H = 1536
    table = torch.randn(5000, H, device=dev, generator=g).bfloat16()
    tok = torch.randint(0, 5000, (4, 97), device=dev, generator=g)
    x, xb, ss = ops.embed_rows(table, tok, with_norm_operands=True)
    assert torch.equal(x, table[tok.reshape(-1)].float()) and torch.equal(xb, table[tok.reshape(-1)])
    ref = (x.double() ** 2).view(-1, H // 128, 128).sum(-1)
    assert ((ss.double() - ref).abs() <= 1e-5 * ref).all()
    assert torch.equal(ops.embed_rows(table, tok), x)
