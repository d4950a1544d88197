#!/bin/bash
# A/B of chain-kernel variants (tools/build_variants.sh) against the in-tree build (first and last, to see the drift)
mkdir -p gpurun_out
V=$PWD/robust-e-nerf_b200/ren_b200/lib/variants
timeout 150 python tools/ab_kernels.py intree_a > gpurun_out/ab3_intree_a.json 2> gpurun_out/ab3_intree_a.err; echo "intree_a rc=$?"
for t in "$@"; do
  REN_B200_LIB=$V/libren_b200_$t.so timeout 100 python tools/ab_kernels.py $t > gpurun_out/ab3_$t.json 2> gpurun_out/ab3_$t.err; echo "$t rc=$?"
done
timeout 100 python tools/ab_kernels.py intree_b > gpurun_out/ab3_intree_b.json 2> gpurun_out/ab3_intree_b.err; echo "intree_b rc=$?"
python - <<'P'
import glob, json
for f in sorted(glob.glob("gpurun_out/ab3_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(d["tag"], "fwd %.3f dgrad %.3f wgrad %.3f" % (d["field_fwd_ms"], d["field_dgrad_ms"], d["wgrad_ms"]),
              "golden %.2e" % d["golden_logI_normwise"], "B37 grad %.2e" % d["B37"]["grad_normwise"],
              "B4096 logI %.2e grad %.2e" % (d["B4096"]["logI_normwise"], d["B4096"]["grad_normwise"]), d["B4096"]["finite"])
    except Exception as e:
        print(f, "unreadable", e)
P
