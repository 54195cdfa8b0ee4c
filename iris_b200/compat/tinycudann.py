"""Stand-in for `tinycudann` as the reference uses it (model/brdf.py:10,239-240): only `NetworkWithInputEncoding(3, 5, enc, net)`
with the HashGrid(32x2, 2^19, base 16, x1.3) + FullyFusedMLP(64, 2 hidden, ReLU) configuration, as a parameter container with one
flat fp32 `params` tensor in tiny-cuda-nn's layout.  Evaluation goes through iris_b200.model.brdf.NGPBRDF (iris_field_forward /
iris_field_backward); calling the module directly raises, loudly, instead of falling back to anything slower."""
from __future__ import annotations

import torch.nn as nn

from ..model.brdf import _FieldParams


class NetworkWithInputEncoding(_FieldParams):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        if n_input_dims != 3 or n_output_dims != 5:
            raise ValueError("only the NGPBRDF configuration (3 -> 5) is implemented")
        e, n = encoding_config, network_config
        want = dict(otype="HashGrid", n_levels=32, n_features_per_level=2, log2_hashmap_size=19, base_resolution=16, per_level_scale=1.3)
        for k, v in want.items():
            if e.get(k) != v:
                raise ValueError("unsupported encoding_config[%r]=%r (kernels are specialised for %r)" % (k, e.get(k), v))
        if n.get("n_neurons") != 64 or n.get("n_hidden_layers") != 2 or n.get("activation") != "ReLU":
            raise ValueError("unsupported network_config %r" % (n,))
        super().__init__(seed)

    def forward(self, x):
        raise RuntimeError("tinycudann shim: evaluate the field through iris_b200.model.brdf.NGPBRDF.forward (CUDA path); "
                           "call iris_b200.compat.install(reference_root) so that model.brdf.NGPBRDF is the CUDA-backed class")


class Module(nn.Module):
    pass
