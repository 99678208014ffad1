"""GPU probe (not a test): %globaltimer stamps of every kernel of one steady-state config-2 window batch
(cond encode -> 50 DDIM steps -> RVQ decode -> 330-d), split into the three phases."""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
from syntalker_b200.pipeline import Window330
from syntalker_b200.vq import RVQVAE
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.set_grad_enabled(False)
L = _lib.lib()
model = MDM(None).load_state_dict(synth.mdm_state_dict("beatx_motionclip", seed=0))
w = ClassifierFreeSampleModel(model)
vqs = [RVQVAE(None, d).load_state_dict(synth.rvq_state_dict(d, seed=0)) for d in synth.PART_DIMS_BEATX]
diff = create_gaussian_diffusion(timestep_respacing="ddim50")
inp = synth.make_inputs(B, seed=1, variant="beatx_motionclip")
d = {k: inp[k].cuda().contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
y = {"scale": torch.ones(1) * 2.0, "style_feature": d["style_feature"]}
win = Window330(w, diff, *vqs, B=B, use_ddim=True)
for _ in range(3):
    win.run_device(d["audio"], d["word"], d["seed"], d["noise"], y=y)
torch.cuda.synchronize()
buf = torch.zeros(120000, dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if os.environ.get("ST_FLUSH") else None
# event time of the same call without stamps, like bench.py takes it (ST_FLUSH=1: a 256 MiB write before it, as bench.py does)
for _ in range(2):
    if flush is not None:
        flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); win.run_device(d["audio"], d["word"], d["seed"], d["noise"], y=y); e1.record()
    torch.cuda.synchronize()
    print(f"window batch by events ({'L2 flushed' if flush is not None else 'warm L2'}): {e0.elapsed_time(e1):.3f} ms")
_lib.check(L.st_debug_trace(buf.data_ptr()))
if flush is not None:
    flush.zero_()
win.run_device(d["audio"], d["word"], d["seed"], d["noise"], y=y)
torch.cuda.synchronize()
_lib.check(L.st_debug_trace(None))
t = buf.cpu().tolist()
n = t[0]
ev = sorted((t[2 * i], t[2 * i + 1]) for i in range(1, n + 1) if t[2 * i + 1] < 20)   # ids >= 20 are in-kernel phase stamps
names = {1: "gemm_tc", 2: "attention", 3: "tokens_in", 4: "step_update", 5: "advance", 6: "split", 7: "layernorm", 8: "gemm_simt", 9: "vq_select", 10: "misc"}
print("kernels stamped:", n, "span us:", (ev[-1][0] - ev[0][0]) / 1e3)
# phases: loop = from first tokens_in-preceding split ... last advance
idx_upd = [i for i, e in enumerate(ev) if e[1] == 4]
idx_tok = [i for i, e in enumerate(ev) if e[1] == 3]
lo, hi = idx_tok[0] - 1, idx_upd[-1]
def show(tag, seg):
    if len(seg) < 2:
        return
    print(f"-- {tag}: {len(seg)} kernels, {(seg[-1][0] - seg[0][0]) / 1e3:.1f} us")
    agg = collections.defaultdict(list)
    for (t0, k0), (t1, _) in zip(seg[:-1], seg[1:]):
        agg[k0].append((t1 - t0) / 1e3)
    for k, v in sorted(agg.items()):
        print(f"   {str(names.get(k, k)):12s} n={len(v):4d} total {sum(v):8.1f} us  mean {sum(v) / len(v):7.2f}  max {max(v):7.2f}")
show("cond encode", ev[:lo + 1])
show("sampling loop", ev[lo:hi + 2])
show("decode + pose", ev[hi + 1:])
print("cond sequence:", " ".join(f"{names.get(k, k)[:5]}:{(b[0] - a) / 1e3:.1f}" for (a, k), b in zip(ev[:lo], ev[1:lo + 1])))
print("decode sequence:", " ".join(f"{names.get(k, k)[:5]}:{(b[0] - a) / 1e3:.1f}" for (a, k), b in zip(ev[hi + 1:-1], ev[hi + 2:])))
