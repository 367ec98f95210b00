"""TEST INFRASTRUCTURE -- the case list of tests/golden/unet_configs.npz, shared by its generator
(oracle/make_golden.py, build container only) and the GPU tests that read the fixture."""
# BASELINE.json configs: config 1 verbatim (R34 b2 @256), the headline net at the headline resolution (R101 @320, batch
# cut to 2 so the CPU reference finishes in seconds), config 5's net and resolution (R152 @512, batch 1)
CONFIG_CASES = (("r34_b2_256", "ResNet34", 34, 2, 256), ("r101_b2_320", "ResNet101", 101, 2, 320),
                ("r152_b1_512", "ResNet152", 152, 1, 512))
CONFIG_GRAD_KEYS = ("encoder.conv1.weight", "encoder.layer1.0.conv1.weight", "encoder.layer2.0.conv2.weight",
                    "encoder.layer3.1.conv1.weight", "encoder.layer4.0.downsample.0.weight", "center.block.0.conv.bias",
                    "dec3.block.1.weight", "dec1.block.1.weight", "dec0.conv.weight", "final.weight", "final.bias")


GRAD_HEAD = 16384   # leading elements (flattened, reference layout) kept of each gradient

# conditioned family (oracle/unet_oracle.py::conditioned_state_dict): same nets / sizes, logits stored at every
# LOGIT_STRIDE-th pixel in both directions (3 covers every 2x2 parity class of the transposed convs)
LOGIT_STRIDE = 3
