"""GPU probe (not a test): what happens between two consecutive trunk layers inside the captured sampling loop.  Every CTA of the last
proj launch (N = 512, K = 512) and of the fc1 launch that follows it (N = 1024, K = 512) records %globaltimer at entry, after the dependency
wait and at exit, plus its SM id (st_debug_probe bit 65536): python tests/boundary_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
B = 32
torch.set_grad_enabled(False)
L = _lib.lib()
dbg = torch.zeros(2048, dtype=torch.int64, device="cuda")
_lib.check(L.st_debug_probe(65536))
_lib.check(L.st_debug_timeline_select(512, 512))
_lib.check(L.st_debug_timeline_select2(1024, 512))
_lib.check(L.st_debug_timeline(dbg.data_ptr()))
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
diff = create_gaussian_diffusion(timestep_respacing="ddim10")
inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
d = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
y = {"audio": d["audio"], "word": d["word"], "seed": d["seed"], "style_feature": d["style_feature"], "scale": torch.ones(1) * 2.0}
for _ in range(4):
    diff.ddim_sample_loop(w, (B, 1536, 1, 32), noise=d["noise"], clip_denoised=False, model_kwargs={"y": y})
torch.cuda.synchronize()
t = dbg.cpu().tolist()
def recs(base, n):
    return [tuple(t[base + 64 + 4 * i: base + 64 + 4 * i + 4]) for i in range(n)]
a, b = recs(0, 128), recs(1024, 128)
a = [r for r in a if r[0]]; b = [r for r in b if r[0]]
t0 = min(r[0] for r in a)
f = lambda x: f"{(x - t0) / 1e3:7.2f}"
print(f"proj: {len(a)} CTAs on {len(set(r[1] for r in a))} SMs; entry {f(min(r[0] for r in a))} .. {f(max(r[0] for r in a))}  dependency {f(min(r[2] for r in a))} .. {f(max(r[2] for r in a))}"
      f"  exit {f(min(r[3] for r in a))} .. {f(max(r[3] for r in a))} us")
print(f"fc1 : {len(b)} CTAs on {len(set(r[1] for r in b))} SMs; entry {f(min(r[0] for r in b))} .. {f(max(r[0] for r in b))}  dependency {f(min(r[2] for r in b))} .. {f(max(r[2] for r in b))}"
      f"  exit {f(min(r[3] for r in b))} .. {f(max(r[3] for r in b))} us")
end_by_sm = {}
for r in a: end_by_sm[r[1]] = max(end_by_sm.get(r[1], 0), r[3])
gaps = sorted((r[0] - end_by_sm[r[1]]) / 1e3 for r in b if r[1] in end_by_sm)
early = sum(1 for r in b if r[1] not in end_by_sm)
print(f"fc1 CTAs on an SM proj did not use: {early}; others enter {gaps[0]:.2f} .. {gaps[len(gaps) // 2]:.2f} (median) .. {gaps[-1]:.2f} us after proj's CTA on that SM has exited")
last_exit = max(r[3] for r in a)
dep = sorted((r[2] - last_exit) / 1e3 for r in b)
print(f"fc1's dependency wait returns {dep[0]:.2f} .. {dep[len(dep) // 2]:.2f} (median) .. {dep[-1]:.2f} us after proj's LAST CTA has exited")
life = sorted((r[3] - r[0]) / 1e3 for r in a)
print(f"proj CTA lifetime {life[0]:.2f} .. {life[len(life) // 2]:.2f} .. {life[-1]:.2f} us; entry -> dependency {sorted((r[2] - r[0]) / 1e3 for r in a)[len(a) // 2]:.2f} us (median)")
