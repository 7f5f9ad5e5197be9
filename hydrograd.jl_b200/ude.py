"""Host-side mirror of the reference's UDE configuration for Manning's n (settings.UDE_settings, create_NN_model):
turns `UDE_choice` + `UDE_NN_config` of the control file (controls/control_settings_2D.jl:86-92, examples/SWE_2D/UDE/*/
run_control.json) into the hg_ude_desc the library takes.  No arithmetic happens here; the network is evaluated and
differentiated by the CUDA kernels of csrc/hg_ude.cu."""
from __future__ import annotations

import numpy as np

from . import _lib as L

CHOICES = {"ManningN_h": 1, "ManningN_h_Umag_ks": 2}
# get_activation, UDE/process_UDE.jl:91-105
ACTIVATIONS = {"relu": 1, "leakyrelu": 2, "sigmoid": 3, "tanh": 4, "softplus": 5}
LAYERNORM = {"none": 0, "cell": 1, "whole": 2}
LUX_EPSILON = float(np.float32(1e-5))


class UDEModel:
    """The network of create_NN_model (UDE/process_UDE.jl:2-65) for UDE_choice "ManningN_h" / "ManningN_h_Umag_ks".

    theta layout (`offsets`) = the ComponentArray of Lux.setup(rng, Chain(...)): per hidden layer the Dense weight
    [width x in] column-major and bias, then the LayerNorm bias and scale; last the output Dense weight and bias.
    layernorm: "whole" = Lux's default dims = Colon() (statistics over the whole width x N array), "cell" = per cell."""

    def __init__(self, UDE_choice: str, UDE_NN_config: dict, layernorm: str = "whole"):
        if UDE_choice not in CHOICES:
            raise ValueError(f"Unknown UDE choice: {UDE_choice}")          # process_ManningN_2D.jl:262 (FlowResistance: not built)
        cfg = UDE_NN_config
        hidden, acts = list(cfg["hidden_layers"]), list(cfg["activations"])
        if len(hidden) != len(acts):
            raise ValueError("The number of hidden layers must match the number of activation functions.")   # process_UDE.jl:19-21
        for a in acts:
            if a not in ACTIVATIONS:
                raise ValueError(f"Unsupported activation function: {a}")                                   # process_UDE.jl:103
        n_in = 1 if UDE_choice == "ManningN_h" else 3
        if int(cfg.get("input_dim", n_in)) != n_in or int(cfg.get("output_dim", 1)) != 1:
            raise ValueError(f"{UDE_choice} needs input_dim = {n_in} and output_dim = 1")
        if not 1 <= len(hidden) <= L.UDE_MAX_HIDDEN or any(not 1 <= w <= L.UDE_MAX_WIDTH for w in hidden):
            raise ValueError(f"hidden_layers must be 1..{L.UDE_MAX_HIDDEN} layers of 1..{L.UDE_MAX_WIDTH} units")
        d = L.UdeDesc()
        d.choice, d.n_hidden, d.layernorm, d.ln_epsilon = CHOICES[UDE_choice], len(hidden), LAYERNORM[layernorm], LUX_EPSILON
        off, n_prev = 0, n_in
        for l, w in enumerate(hidden):
            d.width[l], d.activation[l] = int(w), ACTIVATIONS[acts[l]]
            d.off_weight[l] = off; off += w * n_prev
            d.off_bias[l] = off; off += w
            if layernorm != "none":
                d.off_ln_bias[l] = off; off += w
                d.off_ln_scale[l] = off; off += w
            n_prev = w
        d.off_weight[len(hidden)] = off; off += n_prev
        d.off_bias[len(hidden)] = off; off += 1
        d.n_params = off
        d.h_bounds[:] = [float(x) for x in cfg["h_bounds"]]
        d.umag_bounds[:] = [float(x) for x in cfg.get("Umag_bounds", (0.0, 1.0))]
        d.ks_bounds[:] = [float(x) for x in cfg.get("ks_bounds", (0.0, 1.0))]
        d.output_bounds[:] = [float(x) for x in cfg["output_bounds"]]
        self.desc, self.n_params, self.n_in, self.choice = d, off, n_in, UDE_choice
