"""Registration order of the tensors of one reference ``TransformerModel`` — the order in
which ``state_dict()`` / ``parameters()`` iterate in the reference (src/SEActor.py:170-230:
pos_encoder, transformer_encoder.layers.N [nn.TransformerEncoderLayer members first:
self_attn, linear1, linear2, norm1, norm2; then the SET additions], encoder norm,
rel_encoder, g_encoder, encoder, head).  ``soft_update_network`` zips two modules'
``parameters()`` (common/functional.py:7-10), so the order is part of the contract."""
from __future__ import annotations

from typing import List


def reference_order(kind: str, n_layers: int) -> List[str]:
    out = [f"pos_encoder.embeddings.{i}.weight" for i in range(3)]
    attn = ["in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias",
            "q_proj.weight", "q_proj.bias", "k_proj.weight", "k_proj.bias", "v_proj.weight", "v_proj.bias",
            "vg_proj.weight", "ng_out.weight", "ng_out.bias", "g_out.weight", "g_proj.weight",
            "linear_g1.weight", "linear_g1.bias", "linear_g2.weight", "linear_g2.bias"]
    rest = ["linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
            "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias",
            "g_proj2.weight", "g_proj3.weight", "linear_g1.weight", "linear_g1.bias",
            "linear_g2.weight", "linear_g2.bias", "linear3.weight", "linear3.bias",
            "linear4.weight", "linear4.bias", "linear5.weight"]
    for l in range(n_layers):
        p = f"transformer_encoder.layers.{l}."
        out += [p + "self_attn." + a for a in attn] + [p + r for r in rest]
    out += ["transformer_encoder.norm.weight", "transformer_encoder.norm.bias",
            "transformer_encoder.rel_encoder.weight", "transformer_encoder.rel_encoder.bias",
            "g_encoder.weight", "encoder.weight", "encoder.bias", "gg_proj.weight",
            "linear1_g.weight", "linear1_g.bias", "linear2_g.weight", "linear2_g.bias",
            "linear1_ng.weight", "linear1_ng.bias", "linear2_ng.weight", "linear2_ng.bias"]
    if kind == "critic":
        out += ["decoder_ng.weight", "decoder_ng.bias"]
    else:
        out += ["decoder_g.weight", "linear1_m.weight", "linear1_m.bias", "linear2_m.weight", "linear2_m.bias", "g_proj.weight"]
    return out
