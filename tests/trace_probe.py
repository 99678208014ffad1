"""GPU probe: %globaltimer stamp at the entry of every kernel of one steady-state config-2 sampling loop;
prints start-to-start intervals per kernel kind (= kernel time + gap to the next kernel)."""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion

B = 32
torch.set_grad_enabled(False)
L = _lib.lib()
if os.environ.get("ST_PROBE"):
    _lib.check(L.st_debug_probe(int(os.environ["ST_PROBE"])))
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
diff = create_gaussian_diffusion(timestep_respacing="ddim5")
inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
d = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
y = {"audio": d["audio"], "word": d["word"], "seed": d["seed"], "style_feature": d["style_feature"], "scale": torch.ones(1) * 2.0}
for _ in range(3):
    diff.ddim_sample_loop(w, (B, 1536, 1, 32), noise=d["noise"], clip_denoised=False, model_kwargs={"y": y})
torch.cuda.synchronize()
buf = torch.zeros(120000, dtype=torch.int64, device="cuda")
_lib.check(L.st_debug_trace(buf.data_ptr()))
diff.ddim_sample_loop(w, (B, 1536, 1, 32), noise=d["noise"], clip_denoised=False, model_kwargs={"y": y})
torch.cuda.synchronize()
_lib.check(L.st_debug_trace(None))
t = buf.cpu().tolist()
n = t[0]
ev = sorted((t[2 * i], t[2 * i + 1]) for i in range(1, n + 1))
names = {1: "gemm_tc", 2: "attention", 3: "tokens_in", 4: "step_update", 5: "advance", 6: "split", 7: "layernorm", 8: "gemm_simt"}
phase = [e for e in ev if e[1] >= 20]
ev = [e for e in ev if e[1] < 20]
print("kernels stamped:", n, "span us:", (ev[-1][0] - ev[0][0]) / 1e3)
agg = collections.defaultdict(list)
for (t0, k0), (t1, _k1) in zip(ev[:-1], ev[1:]):
    agg[k0].append((t1 - t0) / 1e3)
for k, v in sorted(agg.items()):
    v2 = sorted(v)
    print(f"{str(names.get(k, k)):12s} n={len(v):4d} start-to-next-start: mean {sum(v) / len(v):7.2f} us  median {v2[len(v2) // 2]:7.2f}  min {v2[0]:7.2f}  max {v2[-1]:7.2f}")
# one diffusion step in order
step = ev[len(ev) // 2: len(ev) // 2 + 50]
print("sequence (kind: us to next):", " ".join(f"{names.get(k, k)[:4]}:{(b[0] - a) / 1e3:.1f}" for (a, k), b in zip(step[:-1], step[1:])))

# CTA (0,0) phases of the trunk kernel inside the chain: 20 entry, 1 dependency resolved, 21 first operands, 22 accumulator ready, 23 done
allev = sorted(phase + [e for e in ev if e[1] == 1])
mid = len(allev) // 2
while allev[mid][1] != 20:
    mid += 1
t0 = allev[mid][0]
lab = {20: "entry", 1: "dep", 21: "ops", 22: "acc", 23: "done"}
print("CTA(0,0) phases, us from an entry:", " ".join(f"{lab[k]}@{(t - t0) / 1e3:.2f}" for t, k in allev[mid:mid + 41]))
