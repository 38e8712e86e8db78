"""GPU parity of the tcgen05 omni-attention kernel (through the C ABI).

Oracle: fp32 eager attention with the closed-form omni mask (oracle/backbone_oracle.py), which is
what the reference computes through transformers' sdpa path with the dense additive mask of
training/prompting_utils.py:975-1020.  Tolerance (bf16 inputs, bf16 probabilities into the second
GEMM, bf16 output, N(0,1) inputs): max-abs <= 2e-2, relative L2 <= 1e-2 -- the same order as torch's
own bf16 SDPA against the fp32 result, which is measured and printed alongside."""
import math

import numpy as np
import pytest
import torch

from oracle import backbone_oracle as BO

pytestmark = pytest.mark.gpu
MAX_ABS_TOL = 2e-2
REL_L2_TOL = 1e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _mk(B, L, Hq, Hkv, D, seed, dev, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    q = (scale * torch.randn(B, L, Hq, D, generator=g)).bfloat16().to(dev)
    k = (scale * torch.randn(B, L, Hkv, D, generator=g)).bfloat16().to(dev)
    v = torch.randn(B, L, Hkv, D, generator=g).bfloat16().to(dev)
    return q, k, v


def _check(out, ref, valid_rows=None, tag=""):
    o, r = out.float(), ref.float()
    if valid_rows is not None:
        o, r = o[valid_rows], r[valid_rows]
    max_abs = (o - r).abs().max().item()
    rel = ((o - r).norm() / r.norm()).item()
    print(f"{tag} max_abs={max_abs:.3e} rel_l2={rel:.3e}")
    assert math.isfinite(max_abs)
    assert max_abs <= MAX_ABS_TOL and rel <= REL_L2_TOL, f"{tag}: max_abs={max_abs} rel_l2={rel}"


def _diagnose(q, k, dump, P, S, scale):
    """Localise a failure: compare the raw score tile S and the probability tile P of the debug dump (filled only by a library built
    with -DUG_ATTN_DUMP=1, `bash tools/build_attn_variants.sh "-DUG_ATTN_DUMP=1"` + UNIGEN_B200_LIB; zeros in the product build)."""
    L = q.shape[1]
    i0, j0 = 0, 0                      # the dump is q tile 0 x its first kv tile (tile 0) of (b=0, h=0)
    rows = min(128, L - i0)
    cols = min(128, L - j0)
    s_ref = (q[0, i0:i0 + rows, 0].float() @ k[0, j0:j0 + cols, 0].float().T)
    s_got = dump[:128 * 128].view(128, 128)[:rows, :cols]
    print("score tile max-abs err:", (s_got - s_ref).abs().max().item(), "ref max", s_ref.abs().max().item())
    print("score tile got[0,:8]", s_got[0, :8].tolist(), "ref[0,:8]", s_ref[0, :8].tolist())
    print("score tile got[1,:4]", s_got[1, :4].tolist(), "ref[1,:4]", s_ref[1, :4].tolist())
    p_got = dump[128 * 128:2 * 128 * 128].view(128, 128)[:rows, :cols]
    print("prob tile got[0,:8]", p_got[0, :8].tolist(), "row sums", p_got.sum(-1)[:4].tolist())


@pytest.mark.parametrize("B,L,P,S", [
    (1, 128, [0], [20]),                       # single tile, no pads
    (2, 256, [0, 37], [60, 100]),              # two tiles, pads inside tile 0
    (2, 395, [100, 0], [137, 137]),            # BASELINE config 1 shape (text 128 -> L = 395)
    (3, 300, [0, 130, 299], [10, 140, 299]),   # pads crossing a tile boundary; all-pad-but-one sequence
    (1, 1163, [57], [137]),                    # BASELINE config 2 length
    (2, 258, [0, 40], [100, 257]),             # L % 256 = 2: GQA-packed tail tile (2 rows x 6 heads in one tile); causal + bidirectional rows in it
    (2, 533, [3, 0], [200, 520]),              # L % 256 = 21: the largest packed tail (126 rows)
    (1, 534, [0], [300]),                      # L % 256 = 22: too many rows to pack, ordinary tail tile
    # BASELINE config 4 (S in {128, 256, 512}, P in {0, S/2}, N in {256, 1024}; L = S + N + 2): the shapes the sweep TIMES
    (2, 386, [0, 64], [128, 128]),             # S = 128, N = 256
    (2, 514, [0, 128], [256, 256]),            # S = 256, N = 256
    (2, 770, [0, 256], [512, 512]),            # S = 512, N = 256
    (2, 1154, [0, 64], [128, 128]),            # S = 128, N = 1024
    (2, 1282, [0, 128], [256, 256]),           # S = 256, N = 1024
    (2, 1538, [0, 256], [512, 512]),           # S = 512, N = 1024
])
def test_omni_attention_matches_fp32_reference(dev, B, L, P, S):
    from ml_unigen_b200 import ops, _lib
    Hq, Hkv, D = 12, 2, 128
    q, k, v = _mk(B, L, Hq, Hkv, D, seed=L + B, dev=dev)
    Pt = torch.tensor(P, dtype=torch.int32, device=dev)
    St = torch.tensor(S, dtype=torch.int32, device=dev)
    dump = torch.zeros(2 * 128 * 128 + 2048, dtype=torch.float32, device=dev)
    out = ops.omni_attention(q, k, v, Pt, St, D ** -0.5, _debug_dump=dump)
    torch.cuda.synchronize()
    ref = BO.attention_ref(q, k, v, Pt, St, D ** -0.5)
    # pad rows (i < P_b) are computed (causal among pads) but unused by the model; they are checked too
    try:
        _check(out, ref, tag=f"L={L}")
    except AssertionError:
        _diagnose(q, k, dump, P, S, D ** -0.5)
        raise
    # what torch's own bf16 SDPA (the reference's kernel) achieves against the same fp32 result
    mask = BO.omni_additive_mask(Pt, St, L, torch.bfloat16)
    sd = torch.nn.functional.scaled_dot_product_attention(
        q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3).repeat_interleave(Hq // Hkv, 1),
        v.permute(0, 2, 1, 3).repeat_interleave(Hq // Hkv, 1), attn_mask=mask, scale=D ** -0.5).permute(0, 2, 1, 3)
    print(f"torch bf16 sdpa vs fp32: max_abs={(sd.float() - ref).abs().max().item():.3e}")


def test_strided_qkv_views_and_other_head_counts(dev):
    """q/k/v as views into one fused [B,L,(Hq+2Hkv)*D] projection output; Hq=4,Hkv=4 (MHA) and 6/1."""
    from ml_unigen_b200 import ops
    D = 128
    for Hq, Hkv in ((4, 4), (6, 1), (12, 2)):
        B, L = 2, 200
        g = torch.Generator().manual_seed(Hq)
        fused = torch.randn(B, L, (Hq + 2 * Hkv) * D, generator=g).bfloat16().to(dev)
        q = fused[..., : Hq * D].view(B, L, Hq, D)
        k = fused[..., Hq * D: (Hq + Hkv) * D].view(B, L, Hkv, D)
        v = fused[..., (Hq + Hkv) * D:].view(B, L, Hkv, D)
        Pt = torch.tensor([3, 0], dtype=torch.int32, device=dev)
        St = torch.tensor([50, 64], dtype=torch.int32, device=dev)
        out = ops.omni_attention(q, k, v, Pt, St, D ** -0.5)
        ref = BO.attention_ref(q, k, v, Pt, St, D ** -0.5)
        _check(out, ref, tag=f"Hq={Hq} Hkv={Hkv}")


def test_large_magnitude_scores_and_full_config_batch(dev):
    """scores with a wide dynamic range (online-softmax rescaling exercised) at B=16, L=1163."""
    from ml_unigen_b200 import ops
    B, L, Hq, Hkv, D = 16, 1163, 12, 2, 128
    q, k, v = _mk(B, L, Hq, Hkv, D, seed=7, dev=dev, scale=2.0)
    g = torch.Generator().manual_seed(1)
    Pt = torch.randint(0, 100, (B,), generator=g).int().to(dev)
    St = torch.full((B,), 137, dtype=torch.int32, device=dev)
    out = ops.omni_attention(q, k, v, Pt, St, D ** -0.5)
    torch.cuda.synchronize()
    for b in (0, 5, 15):        # fp32 reference image by image (memory)
        ref = BO.attention_ref(q[b:b + 1], k[b:b + 1], v[b:b + 1], Pt[b:b + 1], St[b:b + 1], D ** -0.5)
        _check(out[b:b + 1], ref, tag=f"b={b}")
    out2 = ops.omni_attention(q, k, v, Pt, St, D ** -0.5)
    assert torch.equal(out, out2), "attention kernel is not deterministic"


def test_rejects_unsupported_shapes(dev):
    from ml_unigen_b200 import ops, _lib
    q = torch.zeros(1, 16, 2, 64, dtype=torch.bfloat16, device=dev)
    with pytest.raises(_lib.UnigenB200Error):
        ops.omni_attention(q, q, q, torch.zeros(1, dtype=torch.int32, device=dev),
                           torch.zeros(1, dtype=torch.int32, device=dev), 1.0)


def test_rejects_sequences_longer_than_the_work_masks(dev):
    from ml_unigen_b200 import ops, _lib
    q = torch.zeros(1, 4100, 2, 128, dtype=torch.bfloat16, device=dev)
    z = torch.zeros(1, dtype=torch.int32, device=dev)
    with pytest.raises(_lib.UnigenB200Error):
        ops.omni_attention(q, q, q, z, z, 1.0)


def test_longest_supported_sequence(dev):
    """L = 4096 (32 kv tiles: the last bit of the per-tile work masks) and L = 4000 (end-aligned tiles), window pass."""
    from ml_unigen_b200 import ops
    for L, P, S in ((4096, 5, 700), (4000, 200, 300)):
        q, k, v = _mk(1, L, 2, 1, 128, seed=L, dev=dev)
        Pt = torch.tensor([P], dtype=torch.int32, device=dev)
        St = torch.tensor([S], dtype=torch.int32, device=dev)
        s0 = L - 300
        part = ops.omni_attention(q[:, s0:], k, v, Pt, St, 128 ** -0.5, q_offset=s0)
        ref = BO.attention_ref(q, k, v, Pt, St, 128 ** -0.5)[:, s0:]
        _check(part, ref, tag=f"L={L}")


def test_query_row_window_equals_full_attention_rows(dev):
    """ug_omni_attention_rows_bf16 (image-block rows only, the prefix K/V reuse pass) returns exactly the rows the
    full pass returns -- each query row sees the same kv tiles in the same order."""
    from ml_unigen_b200 import ops
    B, L, Hq, Hkv, D = 3, 1163, 12, 2, 128
    q, k, v = _mk(B, L, Hq, Hkv, D, seed=3, dev=dev)
    Pt = torch.tensor([0, 57, 99], dtype=torch.int32, device=dev)
    St = torch.full((B,), 137, dtype=torch.int32, device=dev)
    full = ops.omni_attention(q, k, v, Pt, St, D ** -0.5)
    for s0 in (137, 128, 500, 1162, 1150):      # Lq = 1026 / 1035 (packed tails of 2 / 11 rows), 663, 1 and 13 (a packed tile alone)
        part = ops.omni_attention(q[:, s0:], k, v, Pt, St, D ** -0.5, q_offset=s0)
        assert part.shape == (B, L - s0, Hq, D)
        assert torch.equal(part, full[:, s0:]), f"row window s0={s0} differs from the full pass"


def test_packed_tail_other_group_sizes(dev):
    """GQA-packed tail with 2 and 12 query heads per KV head, against the fp32 reference (window pass)."""
    from ml_unigen_b200 import ops
    D = 128
    for Hq, Hkv, L, s0 in ((4, 2, 600, 304), (12, 1, 700, 434), (4, 4, 300, 40)):
        B = 2
        q, k, v = _mk(B, L, Hq, Hkv, D, seed=L + Hq, dev=dev)
        Pt = torch.tensor([0, 33], dtype=torch.int32, device=dev)
        St = torch.tensor([s0, s0 - 20], dtype=torch.int32, device=dev)
        ref = BO.attention_ref(q, k, v, Pt, St, D ** -0.5)
        part = ops.omni_attention(q[:, s0:], k, v, Pt, St, D ** -0.5, q_offset=s0)
        full = ops.omni_attention(q, k, v, Pt, St, D ** -0.5)
        assert torch.equal(part, full[:, s0:]), (Hq, Hkv, L, s0)
        _check(part, ref[:, s0:], tag=f"packed tail Hq={Hq} Hkv={Hkv}")
