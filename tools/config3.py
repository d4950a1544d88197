"""GPU timing of BASELINE config 3 ("ficus hard": 16384-ray batch x 128 samples, 8x256 MLP, l_diff + l_grad,
C_p and tau trainable with a poor initialisation): 2 value renders + 1 tangent render per step, fwd+bwd+Adam."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "robust-e-nerf_b200"), ROOT):
    sys.path.insert(0, p)
import torch
from ren_b200 import RobustENeRF, synthetic, _lib
from ren_b200.trainer import FlatTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = 128
root = synthetic.write_dataset("/tmp/ren_b200_config3")
cfg = synthetic.default_config(n_samples=N)
cfg["loss"]["weight"]["log_intensity_grad"] = 1e-3
cfg["contrast_threshold"]["freeze"] = False
cfg["refractory_period"]["freeze"] = False
torch.manual_seed(0)
model = RobustENeRF(dataset_directory=root, **cfg).cuda().train()
with torch.no_grad():       # poor initialisation: ratio x1.5, tau x0.5
    ct = model.contrast_threshold.parametrizations.p2n_contrast_threshold_ratio
    ct.original.copy_(ct[0].right_inverse(torch.tensor(1.5)))
    rp = model.refractory_period.parametrizations._refractory_period
    rp.original.copy_(rp[0].right_inverse(model.refractory_period.init_refractory_period.double() * 0.5))
trainer = FlatTrainer(model)
pose_ts = model.trajectory.T_wc_timestamp.cpu()
batches = [synthetic.make_batch(B, pose_ts, rank=0, seed=100 + i) for i in range(3)]
res = [(synthetic.to_device(b, "cuda"), {k: v.cuda() for k, v in j.items()}) for b, j in batches]
lib = _lib.load()
for i in range(3):
    loss = trainer.step(*res[i % 3][:1], i, res[i % 3][1])
torch.cuda.synchronize()
n0 = lib.ren_launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 5
a.record()
for i in range(steps):
    loss = trainer.step(res[i % 3][0], i, res[i % 3][1])
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / steps
print(json.dumps({"workload": f"config3: {B} rays x {N} samples x 3 renders (2 value + 1 tangent), 8x256 MLP, l_diff + l_grad, "
                              "trainable C_p / tau, fwd+bwd+Adam, engine auto (tcgen05)",
                  "ms_per_step": ms, "ray_samples_per_s": 3 * B * N / (ms * 1e-3), "launches_per_step": (lib.ren_launch_count() - n0) / steps,
                  "loss": float(loss), "tau": float(model.refractory_period.refractory_period),
                  "p2n_ratio": float(model.contrast_threshold.p2n_contrast_threshold_ratio)}))
