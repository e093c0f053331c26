"""tcgen05 / TMEM tensor-core contraction (csrc/sg_conv_tc.cu) vs the fp32 FFMA kernel and a torch
fp64 reference of the same op, through sg_debug_conv.  TF32 operands (10-bit mantissa), fp32
accumulate: tolerance 2e-3 relative to the row max (stated; measured ~3e-4)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # rows, cin, N, taps, step, epilogue
    (1000, 32, 512, 5, 1, 1),        # layer 1 forward
    (1300, 512, 512, 5, 2, 1),       # layer 2 forward
    (700, 512, 512, 7, 3, 1),        # layer 3 forward
    (600, 512, 1536, 1, 0, 1),       # layer 5 forward
    (900, 1536, 512, 1, 0, 2),       # layer 5 dgrad (ReLU mask)
    (800, 512, 512, 7, -3, 2),       # layer 3 dgrad
    (640, 512, 32, 5, -1, 3),        # layer 1 dgrad
    (128, 512, 512, 1, 0, 0),        # single tile, bias only
    (40000, 512, 512, 5, 2, 1),      # many tiles per CTA (persistent loop, both TMEM stages, barrier phase wrap)
    (50000, 512, 512, 7, -3, 2),     # long K + many rows: 256 x 256 tile kernel, mask epilogue
    (33000, 1536, 512, 1, 0, 2),     # layer-5 dgrad shape on the 256 x 256 kernel
]


def ref_conv(A, W, bias, rows, N, cin, taps, step, epi, mask, T, tv):
    A64, W64 = A.double(), W.double()
    out = torch.zeros(rows, N, dtype=torch.float64, device=A.device)
    for k in range(taps):
        sh = k * step
        src = torch.zeros_like(A64)
        lo, hi = max(0, -sh), min(rows, rows - sh)
        if hi > lo:
            src[lo:hi] = A64[lo + sh:hi + sh]
        out += src @ W64[k * cin:(k + 1) * cin]
    if epi in (0, 1):
        out += bias.double()
    if epi == 1:
        out = out.clamp_min(0)
    if epi == 2:
        ok = ((torch.arange(rows, device=A.device) % T) < tv).unsqueeze(1)
        out = out * (mask > 0) * ok
    return out


@pytest.mark.parametrize("rows,cin,N,taps,step,epi", SHAPES)
def test_tc_conv_matches_ffma_and_fp64(rows, cin, N, taps, step, epi):
    from speakerguard_b200.engine import Engine, debug_conv
    eng = Engine("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(rows + cin + N)
    A = torch.randn(rows, cin, device="cuda", generator=g)
    W = torch.randn(taps * cin, N, device="cuda", generator=g) / (taps * cin) ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    mask = torch.randn(rows, N, device="cuda", generator=g)
    T, tv = 100, 90
    ref = ref_conv(A, W, bias, rows, N, cin, taps, step, epi, mask, T, tv)
    simt = debug_conv(eng, "fp32", A, W, bias, rows, N, cin, taps, step, epi, mask, T, tv)
    tc = debug_conv(eng, "tf32", A, W, bias, rows, N, cin, taps, step, epi, mask, T, tv)
    torch.cuda.synchronize()
    scale = ref.abs().max()
    e_simt = float((simt.double() - ref).abs().max() / scale)
    e_tc = float((tc.double() - ref).abs().max() / scale)
    print(f"rows={rows} cin={cin} N={N} taps={taps} step={step} epi={epi}: ffma err {e_simt:.2e}  tf32 err {e_tc:.2e}")
    assert e_simt < 1e-5
    assert e_tc < 2e-3


def test_tf32_network_vs_fp32_network():
    """Whole x-vector pass in TF32 tensor-core mode vs the fp32 parity mode (same kernels otherwise).
    Stated tolerances for TF32 mode (<= 2x what was measured on B200: 9e-4 / 6e-4 / 98.6 % / 0.9989): embeddings 2e-3,
    scores 1.5e-3 relative (row max), decisions equal on this data, >= 97.5 % of input-gradient signs equal, gradient
    cosine >= 0.997."""
    from oracle import sg_oracle as O
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    res = {}
    torch.manual_seed(77)
    x = ((torch.rand(6, 1, 32000) * 2 - 1) * 0.5)[:, 0].cuda()
    y = torch.tensor([0, 1, 2, 3, 4, 5]).cuda()
    for prec in ("fp32", "tf32"):
        eng = Engine("cuda:0", precision=prec)
        eng.load_xv(p)
        feat = eng.cmvn(eng.mfcc_fwd(x, _lib.DITHER_PHILOX, None, seed=3, pass_=0, ld=32), ld_out=32)
        emb, ws = eng.embed_fwd(feat)
        scores, dec = eng.score_fwd(emb)
        _, ds = eng.loss(scores, y, make_loss_params("Entropy"))
        dfeat = eng.embed_bwd(eng.score_bwd(emb, ds), ws, 6, feat.shape[1])
        grad = eng.mfcc_bwd(x, eng.cmvn(dfeat, ld_out=32, backward=True), _lib.DITHER_PHILOX, None, seed=3, pass_=0)
        res[prec] = (emb.cpu(), scores.cpu(), dec.cpu(), grad.cpu())
    rel = lambda a, b: float(((a - b).abs().max(1)[0] / b.abs().max(1)[0]).max())
    e_emb, e_sc = rel(res["tf32"][0], res["fp32"][0]), rel(res["tf32"][1], res["fp32"][1])
    g1, g0 = res["tf32"][3], res["fp32"][3]
    sign_agree = float((torch.sign(g1) == torch.sign(g0)).float().mean())
    cos = float((g1 * g0).sum() / (g1.norm() * g0.norm()))
    print(f"tf32 vs fp32: emb {e_emb:.2e} scores {e_sc:.2e} grad sign agreement {sign_agree:.4f} cosine {cos:.5f}")
    assert e_emb < 2e-3 and e_sc < 1.5e-3
    assert torch.equal(res["tf32"][2], res["fp32"][2])
    assert sign_agree > 0.975 and cos > 0.997


BF16_SHAPES = [  # rows, cin, N, taps, step, epilogue, op_bf16, out_bf16
    (1000, 32, 512, 5, 1, 1, False, True),     # layer 1: tf32 operands (fp32 features) -> bf16 activations
    (1300, 512, 512, 5, 2, 1, True, True),     # layers 2..4 forward
    (600, 512, 1536, 1, 0, 1, True, True),     # layer 5 forward
    (900, 1536, 512, 1, 0, 3, True, True),     # dgrad, bf16 in / bf16 out
    (640, 512, 32, 5, -1, 3, True, False),     # layer-1 dgrad: bf16 gradients -> fp32 feature gradient
    (40000, 512, 512, 7, -3, 3, True, True),   # long K, many tiles
]


@pytest.mark.parametrize("rows,cin,N,taps,step,epi,opb,outb", BF16_SHAPES)
def test_tc_conv_bf16(rows, cin, N, taps, step, epi, opb, outb):
    """bf16 operand / output variants vs fp64 on the bf16-rounded inputs.  Tolerance: fp32 accumulate is exact
    to ~1e-5; a bf16 output adds one rounding (2^-9 relative)."""
    from speakerguard_b200.engine import Engine, debug_conv
    eng = Engine("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(rows + cin + N + 1)
    A = torch.randn(rows, cin, device="cuda", generator=g)
    W = torch.randn(taps * cin, N, device="cuda", generator=g) / (taps * cin) ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    Ar = A.to(torch.bfloat16).float() if opb else A
    Wr = W.to(torch.bfloat16).float() if opb else W
    ref = ref_conv(Ar, Wr, bias, rows, N, cin, taps, step, epi, None, 1, 0)
    out = debug_conv(eng, "bf16", A, W, bias, rows, N, cin, taps, step, epi, None, 1, 0, op_bf16=opb, out_bf16=outb)
    torch.cuda.synchronize()
    assert out.dtype == (torch.bfloat16 if outb else torch.float32)
    err = float((out.double() - ref).abs().max() / ref.abs().max())
    print(f"bf16 rows={rows} cin={cin} N={N} taps={taps} opb={opb} outb={outb}: err {err:.2e}")
    assert err < (6e-3 if outb else 2e-3)


def test_bf16_network_vs_fp32_network():
    """Whole x-vector pass in BF16 mode vs fp32 parity mode.  Stated tolerances for BF16 mode (<= 2x what was measured
    on B200: 6e-4 / 4e-4 / 97.3 % / 0.996): embeddings 1.2e-3, scores 8e-4 relative (row max), decisions equal on this data,
    >= 96 % of input-gradient signs equal, gradient cosine >= 0.992."""
    from oracle import sg_oracle as O
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    res = {}
    torch.manual_seed(77)
    x = ((torch.rand(6, 1, 32000) * 2 - 1) * 0.5)[:, 0].cuda()
    y = torch.tensor([0, 1, 2, 3, 4, 5]).cuda()
    for prec in ("fp32", "bf16"):
        eng = Engine("cuda:0", precision=prec)
        eng.load_xv(p)
        feat = eng.cmvn(eng.mfcc_fwd(x, _lib.DITHER_PHILOX, None, seed=3, pass_=0, ld=32), ld_out=32)
        emb, ws = eng.embed_fwd(feat)
        scores, dec = eng.score_fwd(emb)
        _, ds = eng.loss(scores, y, make_loss_params("Entropy"))
        dfeat = eng.embed_bwd(eng.score_bwd(emb, ds), ws, 6, feat.shape[1])
        grad = eng.mfcc_bwd(x, eng.cmvn(dfeat, ld_out=32, backward=True), _lib.DITHER_PHILOX, None, seed=3, pass_=0)
        res[prec] = (emb.cpu(), scores.cpu(), dec.cpu(), grad.cpu())
    rel = lambda a, b: float(((a - b).abs().max(1)[0] / b.abs().max(1)[0]).max())
    e_emb, e_sc = rel(res["bf16"][0], res["fp32"][0]), rel(res["bf16"][1], res["fp32"][1])
    g1, g0 = res["bf16"][3], res["fp32"][3]
    sign_agree = float((torch.sign(g1) == torch.sign(g0)).float().mean())
    cos = float((g1 * g0).sum() / (g1.norm() * g0.norm()))
    print(f"bf16 vs fp32: emb {e_emb:.2e} scores {e_sc:.2e} grad sign agreement {sign_agree:.4f} cosine {cos:.5f}")
    assert e_emb < 1.2e-3 and e_sc < 8e-4
    assert torch.equal(res["bf16"][2], res["fp32"][2])
    assert sign_agree > 0.96 and cos > 0.992


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
def test_fused_pgd_loop_in_tensor_core_modes_matches_fp32_outcome(prec):
    """PGD-5 through sg_pgd_run in the tensor-core modes: same decisions / success as the fp32 parity mode on this
    batch, iterates inside the epsilon ball, >= 95.4 % of the perturbation signs equal after 5 steps and the final loss within
    1.4e-3 (<= 2x what was measured on B200: 97.7 % / 98.4 % and 2.7e-4 / 6.6e-4 for bf16 / tf32)."""
    from oracle import sg_oracle as O
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    torch.manual_seed(5)
    x = ((torch.rand(8, 1, 32000) * 2 - 1) * 0.5)[:, 0].cuda()
    y = torch.randint(0, 10, (8,)).cuda()
    out = {}
    for mode in ("fp32", prec):
        eng = Engine("cuda:0", precision=mode)
        eng.load_xv(p)
        xa = x.clone()
        dec, scores, hist = eng.pgd_run(xa, x, y, max_iter=5, epsilon=0.002, step_size=0.0004, lp=make_loss_params("Entropy"),
                                        dither_mode=_lib.DITHER_PHILOX, seed=11, want_loss_hist=True)
        out[mode] = (xa.cpu(), dec.cpu(), hist.cpu())
    xa0, dec0, h0 = out["fp32"]
    xa1, dec1, h1 = out[prec]
    assert float((xa1 - x.cpu()).abs().max()) <= 0.002 + 1e-7
    assert torch.equal(dec0, dec1)
    agree = float((torch.sign(xa1 - x.cpu()) == torch.sign(xa0 - x.cpu())).float().mean())
    print(f"{prec}: perturbation sign agreement after 5 PGD steps {agree:.4f}; final loss rel diff "
          f"{float((h1[-1] - h0[-1]).abs().max() / h0[-1].abs().max()):.3e}")
    assert agree > 0.954
    assert float((h1[-1] - h0[-1]).abs().max()) < 1.4e-3 * float(h0[-1].abs().max())


def test_pool_adjoint_fused_into_layer5_dgrad_matches_two_kernel_form():
    """SG_OPT_POOL_FUSION (bf16 mode, T >= 128): the statistics-pooling adjoint applied to the staged r5 tiles inside the
    layer-5 dgrad contraction vs the separate pool_bwd pass.  Both round dA5 to bf16 once (the fused form evaluates
    alpha' + beta r with alpha' = alpha - beta mean instead of alpha + beta (r - mean)), so the feature gradients agree
    to bf16 rounding: <= 2e-2 of the per-utterance max (stated), rows beyond the valid frames zero in both."""
    from oracle import sg_oracle as O
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    torch.manual_seed(9)
    # 3 s (T = 300: tiles straddle utterance boundaries at changing offsets) and an odd batch (ragged last tile)
    x = ((torch.rand(7, 1, 48000) * 2 - 1) * 0.5)[:, 0].cuda()
    y = torch.tensor([0, 1, 2, 3, 4, 5, 6]).cuda()
    res = {}
    for fuse in (0, 1):
        eng = Engine("cuda:0", precision="bf16")
        eng.load_xv(p)
        eng.set_option(_lib.OPT_POOL_FUSION, fuse)
        feat = eng.cmvn(eng.mfcc_fwd(x, _lib.DITHER_PHILOX, None, seed=3, pass_=0, ld=32), ld_out=32)
        emb, ws = eng.embed_fwd(feat)
        scores, _ = eng.score_fwd(emb)
        _, ds = eng.loss(scores, y, make_loss_params("Entropy"))
        res[fuse] = eng.embed_bwd(eng.score_bwd(emb, ds), ws, 7, feat.shape[1]).float().cpu()
    g0, g1 = res[0], res[1]
    assert torch.isfinite(g1).all()
    err = float(((g1 - g0).abs().amax((1, 2)) / g0.abs().amax((1, 2))).max())
    cos = float((g1 * g0).sum() / (g1.norm() * g0.norm()))
    print(f"pool fusion vs two-kernel form: max rel diff {err:.2e}, cosine {cos:.6f}")
    assert err < 2e-2 and cos > 0.9995


def test_feat_stash_is_bit_identical_to_recomputing_adjoint():
    """SG_OPT_FEAT_STASH: the fused loop with the forward -> adjoint hand-over gives the same iterates, bit for bit, as the
    adjoint that recomputes the forward (fp32 mode; PGD-3 and EOT size 2)."""
    from oracle import sg_oracle as O
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    torch.manual_seed(21)
    x = ((torch.rand(5, 1, 24000) * 2 - 1) * 0.5)[:, 0].cuda()
    y = torch.tensor([0, 3, 5, 7, 9]).cuda()
    for eot in (1, 2):
        out = []
        for stash in (0, 1):
            eng = Engine("cuda:0", precision="fp32")
            eng.load_xv(p)
            eng.set_option(_lib.OPT_FEAT_STASH, stash)
            xa = x.clone()
            eng.pgd_run(xa, x, y, max_iter=3, epsilon=0.002, step_size=0.0004, lp=make_loss_params("Entropy"),
                        dither_mode=_lib.DITHER_PHILOX, seed=13, eot_size=eot)
            out.append(xa.cpu())
        assert torch.equal(out[0], out[1])


def test_layer1_dgrad_tap_form_matches_long_k_form():
    """SG_OPT_L1_TAP_FORM (bf16 mode): the layer-1 input gradient as a K = 512 contraction into per-tap partial sums + the
    shifted sum over taps vs the K = 2560, N = 32 contraction.  Same bf16 products, fp32 accumulation in a different order:
    <= 1e-5 of the per-utterance max (stated); frames whose taps reach before the first row / across utterances included."""
    from oracle import sg_oracle as O
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    torch.manual_seed(10)
    x = ((torch.rand(5, 1, 32000) * 2 - 1) * 0.5)[:, 0].cuda()
    y = torch.tensor([0, 1, 2, 3, 4]).cuda()
    res = {}
    for tap in (0, 1):
        eng = Engine("cuda:0", precision="bf16")
        eng.load_xv(p)
        eng.set_option(_lib.OPT_L1_TAP_FORM, tap)
        feat = eng.cmvn(eng.mfcc_fwd(x, _lib.DITHER_PHILOX, None, seed=3, pass_=0, ld=32), ld_out=32)
        emb, ws = eng.embed_fwd(feat)
        scores, _ = eng.score_fwd(emb)
        _, ds = eng.loss(scores, y, make_loss_params("Entropy"))
        res[tap] = eng.embed_bwd(eng.score_bwd(emb, ds), ws, 5, feat.shape[1]).float().cpu()
    g0, g1 = res[0], res[1]
    err = float(((g1 - g0).abs().amax((1, 2)) / g0.abs().amax((1, 2))).max())
    print(f"layer-1 dgrad tap form vs long-K form: max rel diff {err:.2e}")
    assert torch.isfinite(g1).all() and err < 1e-5
    assert float(g1[:, :, 30:].abs().max()) == 0.0          # padding columns stay zero


@pytest.mark.parametrize("prec,pool_fusion", [("bf16", 1), ("bf16", 0), ("tf32", 1)])
@pytest.mark.parametrize("B,N", [(7, 48000), (3, 32000), (5, 22400)])
def test_row_compaction_is_bit_identical(prec, pool_fusion, B, N):
    """SG_OPT_ROW_COMPACTION: layer 3 stores only the T - 30 frames that are still valid, layers 4 / 5, the pooling and their
    adjoints run on the compact rows and layer 4's adjoint spreads its result back out.  Every kept value is computed from
    the same operands in the same order, so embeddings and feature gradients must be equal bit for bit (3 s: boxes straddle
    utterance boundaries at changing offsets, ragged last tile; 2 s; 1.4 s: T - 30 = 108 < 128 frames, compaction declines)."""
    from oracle import sg_oracle as O
    from speakerguard_b200 import _lib
    from speakerguard_b200.engine import Engine, make_loss_params
    p = O.make_xv_params(seed=0)
    torch.manual_seed(31)
    x = ((torch.rand(B, 1, N) * 2 - 1) * 0.5)[:, 0].cuda()
    y = (torch.arange(B) % 10).cuda()
    res = {}
    for rc in (0, 1):
        eng = Engine("cuda:0", precision=prec)
        eng.load_xv(p)
        eng.set_option(_lib.OPT_ROW_COMPACTION, rc)
        eng.set_option(_lib.OPT_POOL_FUSION, pool_fusion)
        feat = eng.cmvn(eng.mfcc_fwd(x, _lib.DITHER_PHILOX, None, seed=3, pass_=0, ld=32), ld_out=32)
        emb, ws = eng.embed_fwd(feat)
        scores, _ = eng.score_fwd(emb)
        _, ds = eng.loss(scores, y, make_loss_params("Entropy"))
        g = eng.embed_bwd(eng.score_bwd(emb, ds), ws, B, feat.shape[1])
        res[rc] = (emb.cpu(), g.float().cpu())
    assert torch.isfinite(res[1][1]).all()
    assert torch.equal(res[0][0], res[1][0])
    assert torch.equal(res[0][1], res[1][1])
