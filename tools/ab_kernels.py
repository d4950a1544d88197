"""A/B probe of one build of libren_b200 (selected with REN_B200_LIB): is the tcgen05 engine still on the reference
(rendered log I against the reference-minted golden, parameter gradients and loss against the fp32 engine of the same
library, which the chain-kernel variants do not touch), and what do the three big kernels of a config-2 training
step cost (CUDA events recorded by the library around each launch).  One JSON line on stdout."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import helpers as H
from oracle import synth
from ren_b200 import _lib

lib = _lib.load()
root = synth.write_dataset("/tmp/ren_ab_kernels")
scene = synth.load_scene(root)
out = {"lib": os.environ.get("REN_B200_LIB", "in-tree"), "tag": sys.argv[1] if len(sys.argv) > 1 else ""}
li = lambda r: torch.log(r.double() + 1e-3)

# 1. golden render (reference radiance of the gain-2 network) through the tcgen05 engine
g = H.load_golden("step_8x256_gain2")
rr = g["renders"]["start"]
m = H.build_product(g, root, "cuda", "tc_f16x3")
with torch.no_grad():
    rad, opac, depth, spp = m.nerf(rr["o"].cuda(), rr["d"].cuda(), g["jitter"]["diff_start"].cuda())
nw, ma, _ = H.rel_err(li(rad), li(rr["radiance"]))
out["golden_logI_normwise"], out["golden_logI_maxabs"] = nw, ma

# 2. a training step at a ragged size (last tile partly filled) and at config-2 size: loss / per-ray log I / gradients
#    of the tcgen05 engine against the fp32 engine
N = 128
gg = dict(g)
gg["rcfg"] = dict(g["rcfg"], render_step_size=4.0 / N, far_plane=0.5 + (N + 0.5) * 4.0 / N)
gg["learn_bkgd"] = False
for B in (37, 4096):
    batch, jitter = synth.make_batch(B, scene["T_wc_timestamp"], rank=0, seed=77 + B)
    cb = {k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in batch.items()}
    cj = {k: v.cuda() for k, v in jitter.items()}
    res = {}
    for eng in ("simt_fp32", "tc_f16x3"):
        mm = H.build_product(gg, root, "cuda", eng)
        loss = mm.training_step(cb, 0, cj)
        loss.backward()
        grads = torch.cat([p.grad.reshape(-1).double() for n, p in mm.named_parameters()
                           if n.startswith("nerf.radiance_field.mlp") and p.grad is not None])
        lis = torch.cat((mm.last["start_log_intensity"], mm.last["end_log_intensity"])).double()
        res[eng] = (float(loss), lis, grads)
        if eng == "tc_f16x3" and B == 4096:
            timed = mm
    a, b = res["tc_f16x3"], res["simt_fp32"]
    out[f"B{B}"] = {"loss_rel": abs(a[0] - b[0]) / abs(b[0]),
                    "logI_normwise": float((a[1] - b[1]).norm() / b[1].norm()),
                    "logI_maxabs": float((a[1] - b[1]).abs().max()),
                    "grad_normwise": float((a[2] - b[2]).norm() / b[2].norm()),
                    "grad_cos": float(torch.dot(a[2], b[2]) / (a[2].norm() * b[2].norm())),
                    "finite": bool(torch.isfinite(a[2]).all() and torch.isfinite(a[1]).all())}

# 3. kernel times of the config-2 step (4096 rays x 128 samples x 2 renders)
names = ["field_fwd", "field_dgrad", "wgrad"]
for it in range(3):
    for p in timed.parameters():
        p.grad = None
    timed.training_step(cb, 0, cj).backward()
torch.cuda.synchronize()
lib.ren_timing_enable(1)
for k in range(9):
    lib.ren_timing_collect(k, C.byref(C.c_double()), C.byref(C.c_int64()))
for it in range(10):
    for p in timed.parameters():
        p.grad = None
    timed.training_step(cb, 0, cj).backward()
for k, nme in enumerate(names):
    tot, cnt = C.c_double(), C.c_int64()
    lib.ren_timing_collect(k, C.byref(tot), C.byref(cnt))
    out[nme + "_ms"] = tot.value / max(cnt.value, 1)
lib.ren_timing_enable(0)
print(json.dumps(out), flush=True)
