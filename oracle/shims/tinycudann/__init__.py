"""tinycudann shim -- TEST INFRASTRUCTURE. Pure-PyTorch fp32 restatement of the two tcnn torch
bindings the reference uses (src/UNISLAM.py:242-253, src/networks/decoders.py:50-70), so the
unmodified reference runs on CPU.  PARITY UNPINNED wrt real tcnn (see oracle/grid_ref.py)."""
import torch
import torch.nn as nn

from oracle import grid_ref


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
        super().__init__()
        assert n_input_dims == 3 and encoding_config["otype"] == "HashGrid"
        assert encoding_config["n_features_per_level"] == 2
        self.spec = grid_ref.make_grid_spec(
            int(encoding_config["log2_hashmap_size"]), float(encoding_config["per_level_scale"]),
            int(encoding_config["n_levels"]), int(encoding_config["base_resolution"]))
        self.n_input_dims = 3
        self.n_output_dims = self.spec.n_output_dims
        self.params = nn.Parameter(grid_ref.init_params(self.spec, seed))

    def forward(self, x):
        return grid_ref.encode(self.spec, self.params, x.to(torch.float32).contiguous())


class Network(nn.Module):
    """FullyFusedMLP(ReLU, n_neurons=16, n_hidden_layers=1), no biases, output padded to 16; fp32."""

    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        assert network_config["otype"] == "FullyFusedMLP" and network_config["activation"] == "ReLU"
        assert network_config["n_neurons"] == 16 and network_config["n_hidden_layers"] == 1 and n_input_dims == 32
        self.n_input_dims, self.n_output_dims = n_input_dims, n_output_dims
        self.out_act = {"Tanh": torch.tanh, "Sigmoid": torch.sigmoid, "None": lambda t: t}[network_config["output_activation"]]
        g = torch.Generator().manual_seed(seed)
        xav = lambda o, i: (torch.rand(o, i, generator=g) * 2 - 1) * (6.0 / (i + o)) ** 0.5
        self.params = nn.Parameter(torch.cat([xav(16, 32).reshape(-1), xav(16, 16).reshape(-1)]))

    def forward(self, h):
        W1 = self.params[:512].reshape(16, 32)
        Wo = self.params[512:768].reshape(16, 16)[: self.n_output_dims]
        return self.out_act(torch.relu(h @ W1.t()) @ Wo.t())
