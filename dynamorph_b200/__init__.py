"""dynamorph_b200 -- B200-native (sm_100a) implementation of DynaMorph's VQ-VAE latent-encoding
hot path behind the reference's own Python API:

    dynamorph_b200.HiddenStateExtractor.vq_vae   VQ_VAE, VectorQuantizer, ResidualBlock
    dynamorph_b200.HiddenStateExtractor.vae      VQ_VAE_z16, VQ_VAE_z32
    dynamorph_b200.pipeline.patch_VAE            process_VAE
    dynamorph_b200.pipeline.train_utils          zscore_patch, EarlyStopping
    dynamorph_b200.run_training                  run_one_batch, train
    dynamorph_b200.run_VAE                       `-m process` CLI

All arithmetic runs in hand-written CUDA kernels reached through the C ABI of
libdynamorph_b200.so (include/dynamorph_b200.h).  There is no CPU fallback."""

__version__ = "0.1.0"
