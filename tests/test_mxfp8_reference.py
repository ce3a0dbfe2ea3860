"""The plain-torch model of the MXFP8 quantiser (what the CUDA kernel is checked against bit for bit on a GPU) and the
scale-table layout the block-scaled GEMM consumes: properties that hold on any device."""
import torch

from hypha_b200 import ops


def test_scales_are_the_smallest_power_of_two_that_avoids_saturation():
    torch.manual_seed(0)
    x = (torch.randn(256, 512) * torch.exp(torch.randn(256, 1) * 4)).to(torch.bfloat16)
    x[3, 64:96] = 0
    x[4, :32] = 448.0
    x[5, :32] = 450.0 * torch.ones(32)        # the next bf16 above 448
    q, e = ops.mxfp8_quantize_ref(x)
    assert q.dtype == torch.float8_e4m3fn and e.shape == (256, 16) and e.dtype == torch.uint8
    amax = x.float().reshape(256, 16, 32).abs().amax(-1)
    scale = torch.exp2(e.float() - 127)
    nz = amax > 0
    assert bool(((amax / scale)[nz] <= 448.0).all())                   # nothing saturates ...
    assert bool(((amax / (scale / 2))[nz] > 448.0).all())              # ... and half the scale would
    assert int(e[3, 2]) == 0 and float(q.float()[3, 64:96].abs().max()) == 0.0      # an all-zero block
    assert int(e[4, 0]) == 127 and float(q.float()[4, 0]) == 448.0    # amax == 448 -> scale 1, exact
    assert int(e[5, 0]) == 128                                         # one ulp more -> the next power of two
    back = q.float().reshape(256, 16, 32) * scale[..., None]
    rel = (back.reshape(256, 512) - x.float()).abs() / x.float().abs().clamp_min(1e-30)
    big = x.float().abs() >= amax.repeat_interleave(32, 1) / 8         # the top three binades of every block
    assert float(rel[big & (x != 0)].max()) <= 2 ** -4 + 1e-6          # e4m3: 3 mantissa bits, round to nearest


def test_scale_table_is_the_tcgen05_cp_atom_layout():
    rows, k = 384, 512
    e = (torch.arange(rows * (k // 32)) % 251).to(torch.uint8).reshape(rows, k // 32)
    t = ops.mxfp8_sf_to_table(e)
    assert tuple(t.shape) == ops.mx_sf_shape(rows, k) == (k // 128, rows // 128, 512)
    for r, kb in [(0, 0), (31, 3), (32, 4), (127, 15), (128, 0), (383, 15), (200, 7)]:
        # atom of (K block kb // 4, row block r // 128); inside: 16 bytes per row-in-32, 4 bytes per 32-row group, the
        # four scales of the K block in consecutive bytes = one 32-bit TMEM column of lane r % 32
        assert int(t[kb // 4, r // 128, (r % 32) * 16 + ((r % 128) // 32) * 4 + kb % 4]) == int(e[r, kb])
    # dequantisation through the table inverts the layout
    q = torch.ones(rows, k, dtype=torch.float8_e4m3fn).view(torch.uint8)
    back = ops.mxfp8_dequant_ref(q, t)
    assert torch.equal(back[:, ::32], torch.exp2(e.float() - 127))


def test_round_trip_error_of_a_gaussian_matrix():
    torch.manual_seed(1)
    x = torch.randn(512, 1024).to(torch.bfloat16)
    q, e = ops.mxfp8_quantize_ref(x)
    back = ops.mxfp8_dequant_ref(q.view(torch.uint8), ops.mxfp8_sf_to_table(e))
    rel = ((back - x.float()).norm() / x.float().norm()).item()
    assert 0.01 < rel < 0.04            # ~2^-5 rms for 3 mantissa bits; per-tensor scaling of the same data gives the same
