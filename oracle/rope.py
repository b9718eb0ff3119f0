"""TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product path).

CPU restatement of the caller-side step in front of the attention call in the reference's Wan integration
(reference README.md:301-315):  q_rope = rope_apply(q, grid_sizes, freqs); q_rope = q_rope.bfloat16().

`rope_apply` / `rope_params` are NOT in the reference tree: they are Wan2.1's `wan/modules/model.py` (the model the
reference's README patches; third-party, not vendored, no pinned version in the reference).  Their published
algorithm, restated:

    rope_params(max_seq_len, dim, theta=10000):  freqs = outer(arange(max_seq_len), 1 / theta^(arange(0, dim, 2)/dim))
                                                 (float64), returned as polar(1, freqs)  -> complex128 [max_seq_len, dim/2]
    model:   d = head_dim;  freqs = cat([rope_params(1024, d - 4*(d//6)), rope_params(1024, 2*(d//6)),
                                         rope_params(1024, 2*(d//6))], dim=1)            -> [1024, d/2]
    rope_apply(x, grid_sizes, freqs):   n, c = heads, d/2;  split freqs into [c - 2*(c//3), c//3, c//3] along dim 1;
        for sample i with grid (f, h, w), seq_len = f*h*w:
            x_i = view_as_complex(x[i, :seq_len].to(float64).reshape(seq_len, n, -1, 2))
            freqs_i = cat([freqs[0][:f] broadcast over (f,h,w), freqs[1][:h] ..., freqs[2][:w] ...], -1).reshape(seq_len, 1, -1)
            x_i = view_as_real(x_i * freqs_i).flatten(2);  tokens >= seq_len pass through unchanged
        return stack(...).float()

Parity is pinned only by the algebraic properties tested in tests/test_rope.py (identity at position 0, norm
preservation, composition of rotations) -- there is no reference fixture for this step: "parity unpinned"."""
import torch


def rope_params(max_seq_len, dim, theta=10000.0):
    assert dim % 2 == 0
    freqs = torch.outer(torch.arange(max_seq_len, dtype=torch.float64),
                        1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64) / dim))
    return torch.polar(torch.ones_like(freqs), freqs)


def wan_freqs(head_dim, max_seq_len=1024):
    d = head_dim
    return torch.cat([rope_params(max_seq_len, d - 4 * (d // 6)), rope_params(max_seq_len, 2 * (d // 6)),
                      rope_params(max_seq_len, 2 * (d // 6))], dim=1)


def rope_apply(x, grid_sizes, freqs):
    n, c = x.size(2), x.size(3) // 2
    fr = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    out = []
    for i, (f, h, w) in enumerate(grid_sizes.tolist()):
        seq_len = f * h * w
        x_i = torch.view_as_complex(x[i, :seq_len].to(torch.float64).reshape(seq_len, n, -1, 2))
        freqs_i = torch.cat([fr[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1),
                             fr[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                             fr[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(seq_len, 1, -1)
        x_i = torch.view_as_real(x_i * freqs_i).flatten(2)
        x_i = torch.cat([x_i, x[i, seq_len:].to(torch.float64)])
        out.append(x_i)
    return torch.stack(out).float()
