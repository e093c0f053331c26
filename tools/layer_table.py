#!/usr/bin/env python
"""BENCH-ONLY: per-layer table of the TDNN contractions at the headline shape (B = 1024 utterances x 300 frames):
this repo's conv_tc_kernel (through sg_debug_conv, bf16 operands) against the library kernels torch dispatches for
model/_xv_plda/xvecTDNN.py:49-53 on the same B200 - cuDNN conv1d forward and input-gradient (NCL and channels-last),
and a plain cuBLAS GEMM for the two 1x1 layers.  Times are CUDA-event medians over `--reps` launches after warm-up with
an L2 flush (a 256 MB memset) before every launch.  Writes JSON to stdout / --out.

    python tools/layer_table.py --out profiles/r2_layer_table.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

TDNN = [(32, 512, 5, 1), (512, 512, 5, 2), (512, 512, 7, 3), (512, 512, 1, 1), (512, 1536, 1, 1)]   # padded channel counts


def timed(fn, reps, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1000.0)
    ts.sort()
    return ts[len(ts) // 2]


def raw_conv(eng, A, W, bias, rows, N, cin, taps, tap_step, epilogue, T, t_valid, op_bf16, out_bf16):
    """Pre-convert the operands once and return a closure that only launches conv_tc_kernel (sg_debug_conv)."""
    import ctypes as C
    from speakerguard_b200 import _lib
    Wk = W.t().contiguous()
    if op_bf16:
        A, Wk = A.to(torch.bfloat16).contiguous(), Wk.to(torch.bfloat16).contiguous()
    out = torch.empty(rows, N, device=A.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    keep = (A, W, Wk, bias, out)
    ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())

    def launch():
        rc = eng.lib.sg_debug_conv(eng._h, _lib.PRECISIONS["bf16"], ptr(A), A.shape[1], ptr(W), ptr(Wk), ptr(bias), ptr(out), N, rows,
                                   N, cin, taps, tap_step, epilogue, None, 0, T, t_valid, int(op_bf16), int(out_bf16), eng.stream)
        assert rc == 0, keep and eng.lib.sg_last_error()
    return launch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from speakerguard_b200.engine import Engine
    dev = torch.device("cuda:0")
    eng = Engine(dev, precision="bf16")
    B, T = args.batch, args.frames
    R = B * T
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    torch.backends.cudnn.benchmark = True
    rows = []
    t_in = T
    for li, (ci, co, k, d) in enumerate(TDNN, 1):
        t_out = t_in - (k - 1) * d
        flops = 2.0 * ci * k * co * t_out * B                      # valid frames (what bench.py counts, on padded channels)
        g = torch.Generator(device="cpu").manual_seed(li)
        A = (torch.randn(R, ci, generator=g) * 0.5).to(dev)
        W = (torch.randn(k * ci, co, generator=g) / (k * ci) ** 0.5).to(dev)
        bias = torch.zeros(co, device=dev)
        opb = li > 1                                               # layer 1 reads the fp32 features (tf32 operands)
        ours_f = timed(raw_conv(eng, A, W, bias, R, co, ci, k, d, 1, T, t_out, opb, True), args.reps, flush)
        # dgrad: [R, co] x [k*co, ci], taps walk backwards
        G = (torch.randn(R, co, generator=g) * 0.5).to(dev)
        Wb = (torch.randn(k * co, ci, generator=g) / (k * co) ** 0.5).to(dev)
        ours_b = timed(raw_conv(eng, G, Wb, None, R, ci, co, k, -d, 3, T, t_out, True, li > 1), args.reps, flush)
        # library: cuDNN conv1d on [B, C, T] bf16 (the reference's own layout) and channels-last via conv2d
        x = torch.randn(B, ci, t_in, device=dev, dtype=torch.bfloat16)
        w = torch.randn(co, ci, k, device=dev, dtype=torch.bfloat16)
        bb = torch.zeros(co, device=dev, dtype=torch.bfloat16)
        cudnn_f = timed(lambda: F.relu(F.conv1d(x, w, bb, dilation=d)), args.reps, flush)
        x4 = x.unsqueeze(2).contiguous(memory_format=torch.channels_last)
        w4 = w.unsqueeze(2).contiguous(memory_format=torch.channels_last)
        cudnn_f_cl = timed(lambda: F.relu(F.conv2d(x4, w4, bb, dilation=(1, d))), args.reps, flush)
        gy = torch.randn(B, co, t_out, device=dev, dtype=torch.bfloat16)
        cudnn_b = timed(lambda: torch.nn.grad.conv1d_input(x.shape, w, gy, dilation=d), args.reps, flush)
        gy4 = gy.unsqueeze(2).contiguous(memory_format=torch.channels_last)
        cudnn_b_cl = timed(lambda: torch.nn.grad.conv2d_input(x4.shape, w4, gy4, dilation=(1, d)), args.reps, flush)
        row = {"layer": li, "cin": ci, "cout": co, "taps": k, "dilation": d, "t_in": t_in, "t_out": t_out, "gflop": flops / 1e9,
               "fwd_us": {"conv_tc_kernel(bias+relu+mask)": ours_f, "cudnn_conv1d_ncl(+relu)": cudnn_f,
                          "cudnn_conv2d_channels_last(+relu)": cudnn_f_cl},
               "dgrad_us": {"conv_tc_kernel": ours_b, "cudnn_conv1d_input_ncl": cudnn_b,
                            "cudnn_conv2d_input_channels_last": cudnn_b_cl}}
        if k == 1:
            a2 = torch.randn(R, ci, device=dev, dtype=torch.bfloat16)
            w2 = torch.randn(ci, co, device=dev, dtype=torch.bfloat16)
            row["fwd_us"]["cublas_gemm(no epilogue)"] = timed(lambda: a2 @ w2, args.reps, flush)
            g2 = torch.randn(R, co, device=dev, dtype=torch.bfloat16)
            row["dgrad_us"]["cublas_gemm(no epilogue)"] = timed(lambda: g2 @ w2.t(), args.reps, flush)
        best_f = min(v for kname, v in row["fwd_us"].items() if not kname.startswith("conv_tc"))
        best_b = min(v for kname, v in row["dgrad_us"].items() if not kname.startswith("conv_tc"))
        row["fwd_speedup_vs_best_library"] = best_f / ours_f
        row["dgrad_speedup_vs_best_library"] = best_b / ours_b
        row["fwd_tflops"], row["dgrad_tflops"] = flops / ours_f / 1e6, flops / ours_b / 1e6
        rows.append(row)
        print(json.dumps(row), flush=True)
        t_in = t_out
        del x, w, gy, x4, w4, gy4, A, W, G, Wb
    out = {"shape": {"batch": B, "frames": T}, "precision": "bf16 operands, fp32 accumulate (layer-1 forward: tf32 operands in this repo)",
           "method": "CUDA-event median of %d launches, 256 MB L2 flush before each, cudnn.benchmark=True" % args.reps,
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "layers": rows}
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
