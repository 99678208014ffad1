"""syntalker_b200: B200-native (sm_100a) implementation of SynTalker's diffusion sampling hot path.

Host-side mirrors of the reference's call boundary over libsyntalker_b200.so:
    diffusion.create_gaussian_diffusion / SpacedDiffusion.{p_sample_loop, ddim_sample_loop}
    denoiser.MDM, denoiser_h3d.MDM                      model(x, t, y)
    cfg_sampler.{ClassifierFreeSampleModel, TwoClassifierFreeSampleModel, TwoClassifierFreeSampleModel_Bodypart}
    vq.RVQVAE.latent2origin
    pipeline.Window330 (one native call per window batch, host buffers), pipeline.pose_assemble_330
"""
__version__ = "0.1.0"
