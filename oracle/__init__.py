"""CPU oracle for the SynTalker sampling hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A functional torch-CPU fp32 restatement of the reference's algorithm for the path
`ddim_sample_loop/p_sample_loop -> MDM(+CFG) -> RVQVAE.latent2origin -> 330-d assembly`.
Every function cites the reference file:line it follows. The arithmetic library underneath is the
same one the reference uses (third-party `torch`, unpinned in the reference's requirements.txt:1;
this image has 2.11.0), called on CPU in fp32 exactly where the reference calls it.

Pinning: the reference has NO tests, golden vectors or fixtures for this path.
The oracle is therefore pinned against outputs of the REAL reference modules imported from
/root/reference in the build container: tests/golden/make_golden.py loads the same seeded state
dicts into the reference's `MDM`, `SpacedDiffusion`, CFG wrappers and `RVQVAE`, runs them, and
commits the outputs as tests/golden/*.npz (tests/golden/make_golden_enc.py does the same for RVQVAE.map2latent);
tests/test_oracle_golden.py checks the oracle against them.  oracle/longclip.py restates the trainer's window loop
(not importable here) around those pinned pieces.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package. The product (syntalker_b200/) never does and fails loudly without its CUDA library.
"""
