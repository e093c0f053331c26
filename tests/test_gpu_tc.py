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
