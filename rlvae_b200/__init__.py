"""rlvae_b200 — B200-native metric evaluation + sampling for RlVAE."""
