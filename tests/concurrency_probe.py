"""GPU probe (not a pytest file): does the device have room for a second window batch beside the first?
Every layer of the sampling loop is one wave of 128 CTAs on 148 SMs and ~25 % of a layer's time is hand-over between
dependent kernels, so two INDEPENDENT window batches (two model handles = two workspaces, two streams) may interleave.
Times K window batches of BASELINE config 2 (inputs resident in HBM, CUDA events around all K, 3 warm-up batches):
  serial      one handle, one stream
  two streams handle A on stream 1, handle B on stream 2, K/2 batches each, queued alternately
  B = 64      one handle, one stream, a 64-clip batch (the same total work as two 32-clip batches)
and checks that the concurrent results equal the serial ones bit for bit.
    python tests/concurrency_probe.py [out.json]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from syntalker_b200 import _lib, synth
from syntalker_b200.cfg_sampler import ClassifierFreeSampleModel
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
from syntalker_b200.pipeline import Window330
from syntalker_b200.vq import RVQVAE

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
K = int(os.environ.get("K", "8"))
W_mdm = synth.mdm_state_dict("beatx_motionclip", seed=0)
W_vq = [synth.rvq_state_dict(d, seed=0) for d in synth.PART_DIMS_BEATX]
diff = create_gaussian_diffusion(use_ddim=True)


def make_set(B, seed):
    model = ClassifierFreeSampleModel(MDM(None).load_state_dict(W_mdm))
    vqs = [RVQVAE(None, d).load_state_dict(w) for d, w in zip(synth.PART_DIMS_BEATX, W_vq)]
    inp = synth.make_inputs(B, seed=seed, variant="beatx_motionclip")
    d_in = {k: inp[k].to(dev).contiguous() for k in ("audio", "word", "seed", "noise", "style_feature")}
    y = {"scale": torch.ones(1) * 2.0, "style_feature": d_in["style_feature"]}
    outs = (torch.empty((B, 128, 330), device=dev), torch.empty((B, 128, 3), device=dev), torch.empty((B, 1536, 1, 32), device=dev))
    win = Window330(model, diff, *vqs, B=B, use_ddim=True)
    return {"win": win, "in": d_in, "y": y, "out": outs, "keep": (model, vqs)}


def run(s):
    i = s["in"]
    s["win"].run_device(i["audio"], i["word"], i["seed"], i["noise"], y=s["y"], out=s["out"])


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


res = {}
A, Bs = make_set(32, 1), make_set(32, 2)
for s in (A, Bs):
    for _ in range(3):
        run(s)
torch.cuda.synchronize()
ref_A, ref_B = A["out"][0].clone(), Bs["out"][0].clone()


def serial():
    for _ in range(K):
        run(A)


res["serial_ms_per_batch"] = min(timed(serial) for _ in range(3)) / K
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
main = torch.cuda.current_stream()


def two_streams():
    s1.wait_stream(main)
    s2.wait_stream(main)
    for _ in range(K // 2):
        with torch.cuda.stream(s1):
            run(A)
        with torch.cuda.stream(s2):
            run(Bs)
    main.wait_stream(s1)
    main.wait_stream(s2)


two_streams()
res["two_streams_ms_per_batch"] = min(timed(two_streams) for _ in range(3)) / K
torch.cuda.synchronize()
res["two_streams_bitwise_equal"] = bool(torch.equal(A["out"][0], ref_A) and torch.equal(Bs["out"][0], ref_B))

# more than two in flight (N_SETS=3 python tests/concurrency_probe.py): does a third batch still find gaps?
n_sets = int(os.environ.get("N_SETS", "2"))
if n_sets > 2:
    sets = [A, Bs] + [make_set(32, 10 + i) for i in range(n_sets - 2)]
    for s in sets[2:]:
        for _ in range(3):
            run(s)
    strs = [torch.cuda.Stream() for _ in sets]

    def n_streams():
        for st in strs:
            st.wait_stream(main)
        for i in range(K * n_sets // 2):
            with torch.cuda.stream(strs[i % n_sets]):
                run(sets[i % n_sets])
        for st in strs:
            main.wait_stream(st)

    n_streams()
    res[f"{n_sets}_streams_ms_per_batch"] = min(timed(n_streams) for _ in range(3)) / (K * n_sets // 2)
    del sets
del Bs
C64 = make_set(64, 3)
for _ in range(3):
    run(C64)
res["b64_ms_per_32_clips"] = min(timed(lambda: [run(C64) for _ in range(K // 2)]) for _ in range(3)) / K
for k in list(res):
    if k.endswith("_ms_per_batch") or k.endswith("_clips"):
        res[k.replace("_ms_per_batch", "_frames_per_s").replace("_ms_per_32_clips", "_frames_per_s")] = 32 * 128 / (res[k] / 1e3)
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
