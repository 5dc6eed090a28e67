"""Kernel-level parity of the tensor-core contractions (conv_tma.cu, conv_tc.cu, conv_tc_wgrad.cu), through the
C ABI: a 1x1 convolution feeding a kxk convolution; the forward output, the input adjoint (dgrad) and the weight
gradient (wgrad) of the second convolution are compared with an fp64 torch evaluation of the same fp32 inputs.

The shapes are the three DenseNet3 block geometries (32x32, 16x16, 8x8 images; the 8x8 tile holds two images),
ragged channel counts (chunk tails, 1 k-step k-blocks) and both kernel sizes.  Tolerance: fp32-accurate,
1e-5 relative L2 (3xTF32 measured 2e-7 .. 4e-7), i.e. 10x inside the 1e-4 budget of BASELINE.json.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,cin,cout,batch,k", [
    (32, 48, 12, 4, 3),      # DenseNet3 block 1 bottleneck 3x3 (two channel chunks: 32 + 16)
    (32, 60, 48, 4, 1),      # ... and its 1x1 (ragged chunk tail, BN = 48)
    (16, 40, 12, 8, 3),      # block 2; 40 channels: the tail chunk is a single k-step
    (8, 132, 48, 32, 1),     # block 3: tile = two 8x8 images, five chunks
    (8, 48, 12, 32, 3),
    (32, 12, 24, 3, 3),      # batch 3: ragged last pixel tile
])
def test_tensor_core_contractions_match_fp64(W, cin, cout, batch, k):
    import tma_probe
    errs = tma_probe.probe(W, cin, cout, batch, k, verbose=False)
    for (mode, what), e in errs.items():
        assert e < 1e-5, "mode %d %s: relative L2 error %.3e" % (mode, what, e)
    # the tensor-core path must not be (much) less accurate than the fp32 CUDA-core path
    for what in ("fwd", "dgrad", "wgrad"):
        assert errs[(2, what)] < 4 * errs[(0, what)] + 5e-7, (what, errs)


@pytest.mark.parametrize("W,cin,cout,batch,k,stride,pad", [
    (14, 64, 96, 4, 3, 1, 1),     # VGG16 conv5_x / DenseNet121 block 3 geometry: rows of 14 floats are not 16-byte aligned
    (7, 128, 32, 8, 3, 1, 1),     # DenseNet121 block 4; 49-pixel images: k-groups straddle rows and samples
    (14, 40, 24, 5, 5, 1, 2),     # 5 x 5 kernel: horizontal shifts of two pixels
    (15, 24, 40, 6, 3, 2, 1),     # stride 2, odd width
    (12, 32, 16, 7, 3, 1, 0),     # no padding: 12 -> 10 wide output
])
def test_weight_gradient_scalar_gather_variant_matches_fp64(W, cin, cout, batch, k, stride, pad):
    """conv_tc_wgrad_kernel<BN, VEC=false>: geometries the 128-bit transform cannot take (widths that are not a multiple
    of 4, strides, wide kernels).  Mode 2 forces the tensor-core kernels wherever they are legal."""
    import tma_probe
    errs = tma_probe.probe(W, cin, cout, batch, k, verbose=False, stride=stride, pad=pad)
    for (mode, what), e in errs.items():
        assert e < 1e-5, "mode %d %s: relative L2 error %.3e" % (mode, what, e)
    # few pixels, many channels (7 x 7 maps): 1.1e-6 measured for the 3xTF32 contraction against 2e-7 for fp32 FMAs; the
    # order of the fp32 atomics moves both by ~10 % from run to run
    assert errs[(2, "wgrad")] < 4 * errs[(0, "wgrad")] + 2e-6, errs


@pytest.mark.parametrize("W,cin,cout,batch,k", [
    (56, 64, 64, 2, 3),      # VGG16 / DenseNet121 widths: flattened-plane boxes, two channel groups of 32 x 3 taps
    (28, 96, 128, 3, 1),     # 1 x 1, 784-pixel planes: the last k-block of every image is ragged (zero-filled by TMA)
    (28, 100, 136, 2, 3),    # ragged channel groups (3 x 34 rows per tile), two column tiles of 128
    (12, 20, 24, 5, 3),      # 144-pixel planes, rows of 12: horizontal neighbours across row ends are masked
    (40, 16, 200, 1, 3),     # few shifted channels (operands exchanged), wide other side
])
def test_weight_gradient_flat_plane_variant_matches_fp64(W, cin, cout, batch, k):
    """conv_wgrad_tma_kernel<BN, FLAT=true>: k-blocks of 32 consecutive pixels of the flattened image plane."""
    import tma_probe
    errs = tma_probe.probe(W, cin, cout, batch, k, verbose=False)
    for (mode, what), e in errs.items():
        assert e < 1e-5, "mode %d %s: relative L2 error %.3e" % (mode, what, e)
    assert errs[(2, "wgrad")] < 4 * errs[(0, "wgrad")] + 5e-7, errs
