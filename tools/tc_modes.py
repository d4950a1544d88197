"""Speed and precision of the tcgen05 operand-precision modes on BASELINE config 2 (4096 rays x 128 samples x 2
renders, 8x256): per-kernel CUDA-event times of a training step, and per-ray log I against the fp32 engine and
(on a ray subset) the torch-CPU oracle, for the default-init network, the weights-x2 network and a network
trained for some steps."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import helpers as H
from oracle import restatement as R, synth
from ren_b200 import _lib
from ren_b200.trainer import FlatTrainer

lib = _lib.load()
root = synth.write_dataset("/tmp/ren_tc_modes")
scene = synth.load_scene(root)
N, B = 128, 4096
step = 4.0 / N
names = ["field_fwd", "field_dgrad", "wgrad"]
out = {}


def golden(gain):
    g = dict(H.load_golden("step_8x256_gain2" if gain else "step_8x256"))
    g["rcfg"] = dict(g["rcfg"], render_step_size=step, far_plane=0.5 + (N + 0.5) * step)
    g["learn_bkgd"] = False
    return g


batch, jitter = synth.make_batch(B, scene["T_wc_timestamp"], rank=0, seed=1234)
cb = {k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in batch.items()}
cj = {k: v.cuda() for k, v in jitter.items()}

# a trained network: 150 Adam steps (lr 5e-4) on 4 rotating batches with the full-precision engine
g2 = golden(True)
mt = H.build_product(g2, root, "cuda", "auto")
mt.hparams["optimizer"]["lr"]["default"] = 5e-4
tr = FlatTrainer(mt)
bs = [synth.make_batch(B, scene["T_wc_timestamp"], rank=0, seed=900 + i) for i in range(4)]
for i in range(150):
    b, j = bs[i % 4]
    l = tr.step({k: {kk: vv.cuda() for kk, vv in v.items()} for k, v in b.items()}, i, {k: v.cuda() for k, v in j.items()})
trained = {k: v.detach().clone() for k, v in mt.state_dict().items() if k.startswith("nerf.radiance_field.mlp.")}
print("trained loss", float(l), flush=True)
del mt, tr
torch.cuda.empty_cache()

for net in ("default", "gain2", "trained"):
    g = golden(net != "default")
    ref_li = None
    for engine in ("simt_fp32", "tc_f16x3", "tc_f16a", "tc_f16"):
        m = H.build_product(g, root, "cuda", engine)
        if net == "trained":
            m.load_state_dict(trained, strict=False)
        rec = {}
        if engine != "simt_fp32":
            for it in range(3):
                for p in m.parameters():
                    p.grad = None
                m.training_step(cb, 0, cj).backward()
            torch.cuda.synchronize()
            lib.ren_timing_enable(1)
            for k in range(9):
                lib.ren_timing_collect(k, C.byref(C.c_double()), C.byref(C.c_int64()))
            for it in range(5):
                for p in m.parameters():
                    p.grad = None
                m.training_step(cb, 0, cj).backward()
            for k, nme in enumerate(names):
                tot, cnt = C.c_double(), C.c_int64()
                lib.ren_timing_collect(k, C.byref(tot), C.byref(cnt))
                rec[nme + "_ms"] = tot.value / max(cnt.value, 1)
            lib.ren_timing_enable(0)
        with torch.no_grad():
            loss = m.training_step(cb, 0, cj)
        li = torch.cat((m.last["start_log_intensity"], m.last["end_log_intensity"])).double().cpu()
        if engine == "simt_fp32":
            ref_li = li
            # oracle on the first 256 rays
            sub = {k: {kk: vv[:, :256] for kk, vv in v.items()} for k, v in batch.items()}
            sj = {k: v[:256] for k, v in jitter.items()}
            fcfg, rcfg, scfg = H.golden_configs(g)
            params = {k[len("nerf.radiance_field."):]: v.detach().cpu() for k, v in m.state_dict().items()
                      if k.startswith("nerf.radiance_field.mlp.")}
            tau_max = scene["max_refractory_period"]
            with torch.no_grad():
                o = R.training_step(params, fcfg, rcfg, scfg, scene, sub, sj, torch.log(torch.expm1(torch.tensor(1.0))),
                                    tau_max * torch.logit(scene["refractory_period"].double() / tau_max))
            orc = torch.cat((o["start_log_intensity"], o["end_log_intensity"])).double()
        sel = torch.cat((li[:256], li[B:B + 256]))
        rec.update(loss=float(loss), logI_std=float(li.std()),
                   vs_fp32_engine_normwise=float((li - ref_li).norm() / ref_li.norm()),
                   vs_fp32_engine_maxabs=float((li - ref_li).abs().max()),
                   vs_oracle_normwise=float((sel - orc).norm() / orc.norm()),
                   vs_oracle_maxabs=float((sel - orc).abs().max()))
        out[f"{net}/{engine}"] = rec
        print(net, engine, json.dumps(rec), flush=True)
        del m
        torch.cuda.empty_cache()
print(json.dumps(out))
