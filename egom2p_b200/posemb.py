"""Fixed sin-cos positional tables (reference: egom2p/models/egom2p_utils.py:32-44 1-D, :63-86 3-D).
Built once at module construction on the host (they are persistent buffers, not hot-path work); the op order follows
the reference so the tables are bit-identical (tests/test_posemb.py checks SHA-256 digests of the reference's)."""
import torch


def build_1d_sincos_posemb(max_len: int, embed_dim: int = 1024, temperature: float = 10000.0) -> torch.Tensor:
    assert embed_dim % 2 == 0, "Embed dimension must be divisible by 2 for 1D sin-cos position embedding"
    positions = torch.arange(max_len, dtype=torch.float32)
    half = embed_dim // 2
    freq = 1.0 / (temperature ** (torch.arange(half, dtype=torch.float32) / half))
    angles = torch.einsum("n,d->nd", [positions, freq])
    return torch.cat([torch.sin(angles), torch.cos(angles)], dim=1).unsqueeze(0)  # (1, N, D)


def build_3d_sincos_posemb(t: int, h: int, w: int, embed_dim: int = 1024, temperature: float = 10000.0) -> torch.Tensor:
    assert embed_dim % 6 == 0, "Embed dimension must be divisible by 6 for 3D sin-cos position embedding"
    ch = int(embed_dim // 6 * 2)
    inv_freq = 1.0 / (temperature ** (torch.arange(0, ch, 2).float() / ch))

    def axis_table(n: int) -> torch.Tensor:  # (n, ch), sin/cos interleaved
        ang = torch.einsum("i,j->ij", torch.arange(n, dtype=torch.float32), inv_freq)
        return torch.flatten(torch.stack((ang.sin(), ang.cos()), dim=-1), -2, -1)

    table = torch.zeros((1, t, h, w, ch * 3), dtype=torch.float32)
    table[..., :ch] = axis_table(t)[:, None, None, :]
    table[..., ch:2 * ch] = axis_table(h)[None, :, None, :]
    table[..., 2 * ch:] = axis_table(w)[None, None, :, :]
    return table.reshape(1, t * h * w, embed_dim)
