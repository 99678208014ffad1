"""GPU debug probe: 1000-step p_sample_loop under graph / PDL toggles, each in its own process."""
import os, subprocess, sys
CHILD = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
from syntalker_b200 import _lib, synth
from syntalker_b200.denoiser import MDM
from syntalker_b200.diffusion import create_gaussian_diffusion
torch.set_grad_enabled(False)
L = _lib.lib()
L.st_set_graphs(int(os.environ["G"])); L.st_set_pdl(int(os.environ["P"]))
m = MDM(None).load_state_dict(synth.mdm_state_dict("beatx", seed=0))
S = os.environ["S"]
d = create_gaussian_diffusion(timestep_respacing=[int(S)] if S != "1000" else None)
inp = synth.make_inputs(2, seed=51)
y = {k: inp[k].cuda() for k in ("audio", "word", "seed")}
d6 = create_gaussian_diffusion(timestep_respacing=[6])
if os.environ.get("PRE", "0") == "1":
    for i in range(3):
        d6.p_sample_loop(m, (2, 1536, 1, 32), noise=inp["noise"].cuda(), clip_denoised=False, model_kwargs={"y": y})
outs = []
for i in range(3):
    torch.manual_seed(5)
    outs.append(d.p_sample_loop(m, (2, 1536, 1, 32), noise=inp["noise"].cuda(), clip_denoised=False, model_kwargs={"y": y}))
    if os.environ.get("SYNC", "1") == "1":
        torch.cuda.synchronize()
torch.cuda.synchronize()
print("ok", [float(o.abs().max()) for o in outs], flush=True)
'''
for S, G, P, PRE, SYNC in (("1000", "1", "1", "1", "0"), ("1000", "1", "1", "0", "0"), ("1000", "0", "1", "0", "0"), ("1000", "1", "0", "1", "0"),
                          ("200", "1", "1", "1", "0"), ("1000", "1", "1", "1", "1")):
    if True:
        env = dict(os.environ, G=G, P=P, S=S, PRE=PRE, SYNC=SYNC)
        r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
        tail = (r.stdout + r.stderr).strip().splitlines()[-3:]
        print(f"S={S} graphs={G} pdl={P} pre={PRE} sync={SYNC}: rc={r.returncode} | " + " | ".join(t[:160] for t in tail), flush=True)
