"""Discrete-event model of the chain kernel's per-layer schedule (DESIGN.md sections 4.1 / 8.1), no GPU needed.

One 128-sample tile is in flight per SM.  A 256-wide layer's output columns are split into chunks (today two halves).
Resources: ONE tensor pipe that executes MMA blocks in program order, ONE epilogue unit (the 16 epilogue warps work
on one chunk at a time, in order).  Program order of a layer with chunks 0..n-1 of column fractions f_i:
    for k < n-1:  block [K-chunk k x all N]            (f_k M cycles)        needs epilogue chunk k of the layer before
    for i < n:    block [K-chunk n-1 x N-chunk i]      (f_{n-1} f_i M)       needs epilogue chunk n-1 of the layer before;
                                                                             completes accumulator chunk i
    epilogue chunk i (f_i E cycles) needs accumulator chunk i and epilogue chunk i-1.
Hand-offs: `h_acc` cycles from the last MMA of an accumulator chunk to the first epilogue instruction on it (commit ->
mbarrier -> waiting warps -> tcgen05.ld), `h_a` from the end of an epilogue chunk to the first dependent MMA (tcgen05.st
-> wait::st -> 16 arrivals -> issuer -> tensor pipe).

    python tools/chain_model.py            # the table quoted in DESIGN.md section 8.1
"""
import argparse


def period(fracs, E, M, h_acc, h_a, layers=40):
    n = len(fracs)
    pipe_free = 0.0                       # tensor pipe: in-order
    epi_free = 0.0                        # epilogue unit: in-order
    epi_done_prev = [0.0] * n             # epilogue chunk completion times of the previous layer
    starts = []
    for _ in range(layers):
        acc_ready = [0.0] * n
        # first n-1 K-chunks: all N at once
        for k in range(n - 1):
            t0 = max(pipe_free, epi_done_prev[k] + h_a)
            pipe_free = t0 + fracs[k] * M
        # last K-chunk per N-chunk; its end completes accumulator chunk i
        for i in range(n):
            t0 = max(pipe_free, epi_done_prev[n - 1] + h_a)
            pipe_free = t0 + fracs[n - 1] * fracs[i] * M
            acc_ready[i] = pipe_free
        epi_done = [0.0] * n
        for i in range(n):
            t0 = max(epi_free, acc_ready[i] + h_acc)
            if i == 0:
                starts.append(t0)
            epi_free = t0 + fracs[i] * E
            epi_done[i] = epi_free
        epi_done_prev = epi_done
    return starts[-1] - starts[-2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--E", type=float, default=6200.0, help="epilogue cycles per 256-wide layer (16 warps)")
    ap.add_argument("--M", type=float, default=6500.0, help="tensor-pipe cycles per 256-wide layer (48 K-step MMAs x 3)")
    ap.add_argument("--h-acc", type=float, default=700.0)
    ap.add_argument("--h-a", type=float, default=700.0)
    a = ap.parse_args()
    nominal = 6144.0                      # 16 K steps x 3 MMAs x 128 cycles (M=128, N=256, K=16, fp16)
    rows = [
        ("two halves (the kernel today)", [0.5, 0.5], a.M),
        ("two halves, MMAs issue-bound by the lane-0 issuer (M = 8.4 k: round 1 / early round 2)", [0.5, 0.5], 8400.0),
        ("64 | 192 columns", [0.25, 0.75], a.M),
        ("96 | 160 columns", [0.375, 0.625], a.M),
        ("four quarters", [0.25] * 4, a.M),
        ("32 | 224 columns", [0.125, 0.875], a.M),
    ]
    print(f"E = {a.E:.0f}, M = {a.M:.0f}, hand-offs {a.h_acc:.0f} + {a.h_a:.0f} cycles; nominal tensor time {nominal:.0f} per layer")
    print(f"{'split':<90s} {'period':>8s} {'tensor pipe':>12s}")
    for name, fr, M in rows:
        p = period(fr, a.E, M, a.h_acc, a.h_a)
        print(f"{name:<90s} {p:8.0f} {100 * min(nominal, M) * (5650.0 / 6144.0) / p:11.1f}%")
    print("(tensor pipe = measured active cycles per layer today, 5.65 k = 60 % of 9.4 k, over the modelled period)")


if __name__ == "__main__":
    main()
