"""Synthetic Poisson-Gaussian Bayer data for benchmarks and profiling (no checkpoints / datasets are reachable offline).

Follows the reference's own noise synthesis (data_process/yond_datasets.py:664-682, :720):
    log K ~ U(-2.5, 3.5);  log sigma ~ N((0.85187 +- 0.2) log K + (0.67991 +- 1), 0.02921);
    y = Poisson(x / beta1) * beta1 + N(0, beta2)   in normalised units, beta1 = K / scale, beta2 = sigma / scale.
NumPy only; nothing here is on the device path."""
import numpy as np


def sample_noise_params(rng, logk_min=-2.5):
    """(K, sigma) drawn like the reference, redrawn until sigma/K is inside the BiasLUT's range (< 10 e-)."""
    while True:
        logK = rng.uniform(logk_min, 3.5)
        mu = (0.85187 + rng.uniform(-0.2, 0.2)) * logK + (0.67991 + rng.uniform(-1, 1))
        K = float(np.exp(logK))
        sigma = float(np.exp(rng.normal(mu, 0.02921)))
        if sigma / K < 9.5:
            return K, sigma


def clean_smooth(rng, H, W):
    """Smooth clean Bayer field in [0.03, 0.95]: low-frequency shading, a per-frame level and a mild per-CFA-site cast,
    no edges, so that (as in the flat regions of photographs) the 29x29 local statistics are dominated by the noise."""
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    level = rng.uniform(0.08, 0.7)
    amp = rng.uniform(0.2, 1.0) * 0.04 * (H / 256.0)
    f = rng.uniform(0.3, 1.2, size=3) * 2 * np.pi
    ph = rng.uniform(0, 2 * np.pi, size=3)
    field = (np.sin(f[0] * yy / H + ph[0]) * np.cos(f[1] * xx / W + ph[1]) + 0.5 * np.sin(f[2] * (xx / W + yy / H) + ph[2])) / 1.5
    img = level + amp * field
    cast = rng.uniform(0.85, 1.0, size=(2, 2)).astype(np.float32)
    img = img * np.tile(cast, (H // 2, W // 2))
    return np.clip(img, 0.03, 0.95).astype(np.float32)


def noisy(rng, clean, K, sigma, scale=959.0, clip=True):
    """Poisson-Gaussian observation of `clean` (normalised units)."""
    b1, s2 = K / scale, sigma / scale
    out = rng.poisson(clean / b1).astype(np.float32) * b1 + rng.normal(0, s2, clean.shape).astype(np.float32)
    if clip:
        out = np.clip(out, 0, 1)
    return out.astype(np.float32)


def random_init_state_dict(arch, seed=0):
    """Random-init weights of a yml-configured arch, exactly as the reference driver produces them when no checkpoint is
    loaded: construct the module, then `initialize_weights` (kaiming-normal convs) under `torch.manual_seed(seed)`."""
    import torch

    from . import archs
    torch.manual_seed(seed)
    net = getattr(archs, arch["name"])(arch)
    archs.initialize_weights(net)
    return {k: v.detach().clone() for k, v in net.state_dict().items()}


def smoother_state_dict(arch, alpha=1.0):
    """Reference-layout weights that make either architecture a 3x3 mean filter, out = (1-alpha) x + alpha blur3(x), using
    only the first conv -> skip -> last block -> 1x1 head (every other tensor zero)."""
    import torch
    sd = {k: torch.zeros_like(v) for k, v in random_init_state_dict(arch, seed=0).items()}
    hp = torch.full((3, 3), 1.0 / 9.0)
    hp[1, 1] -= 1.0
    if arch["name"] == "UNetSeeInDark":
        g = 1.0 / (1.0 + 0.2)
        for c in range(4):
            sd["conv1_1.weight"][c, c] = hp
            sd["conv1_1.weight"][c + 4, c] = -hp
        for name, off in (("conv1_2", 0), ("conv9_1", 32), ("conv9_2", 0)):
            for c in range(4):
                w = sd[name + ".weight"]
                w[c, off + c, 1, 1], w[c, off + c + 4, 1, 1] = g, -g
                w[c + 4, off + c, 1, 1], w[c + 4, off + c + 4, 1, 1] = -g, g
        for c in range(4):
            sd["conv10_1.weight"][c, c, 0, 0], sd["conv10_1.weight"][c, c + 4, 0, 0] = alpha * g, -alpha * g
    else:
        g = 1.0 / (1.0 + 0.01)
        for c in range(4):
            sd["conv_in.weight"][c, c] = hp
            sd["conv_in.weight"][c + 4, c] = -hp
        for c in range(32):
            sd["conv9.short_cut.0.weight"][c, 32 + c, 0, 0] = 1.0
        for c in range(4):
            sd["conv10.weight"][c, c, 0, 0], sd["conv10.weight"][c, c + 4, 0, 0] = alpha * g, -alpha * g
    return sd


def bench_state_dict(arch, seed=0, eps=0.25):
    """Weights for throughput runs: no checkpoint is reachable offline, and with the reference's plain random init the
    round-2 estimate is degenerate (beta1 < 0), so the reference's guard (YOND_SIDD.py:445-447) would skip the second
    network pass that the shipped pipeline runs.  A mild smoother plus `eps` x the reference's random init (every tensor
    dense) behaves like a weak denoiser: the guard passes and both rounds execute, in the GPU arm and in the CPU arm alike.
    FLOPs and memory traffic do not depend on the weight values."""
    rnd, sm = random_init_state_dict(arch, seed=seed), smoother_state_dict(arch)
    return {k: sm[k] + eps * rnd[k] for k in rnd}
