"""CPU oracle for ``factory/LstmDV.py:4-24``.  Test infrastructure."""
import torch

from .layers import lstm_stack, cast_state_dict


@torch.no_grad()
def lstmdv_forward(sd, x, dtype=torch.float32, lstm_impl="aten", taps=None):
    """3x LSTM(80->768) -> last step -> Linear(768->256) -> e/||e||_2 (no epsilon).  x (B,T,80) -> (B,256)."""
    sd = cast_state_dict(sd, dtype)
    out = lstm_stack(sd, "lstm", x.to(dtype), num_layers=3, impl=lstm_impl)   # LstmDV.py:20
    last = out[:, -1, :]                                                       # :21
    if taps is not None:
        taps["h_last"] = last
    e = last @ sd["embedding.weight"].t() + sd["embedding.bias"]              # :21
    return e / e.norm(p=2, dim=-1, keepdim=True)                               # :22-24


@torch.no_grad()
def lstmdv_twin_forward(sd, x, dtype=torch.float32, lstm_impl="aten"):
    """The classifier twin (make_data/factory/LstmDV.py:18-25): returns (predictions, d_vec) where
    ``predictions = output(embeds)`` uses the UN-normalised embedding (:24) and ``d_vec = embeds / ||embeds||`` (:22-23)."""
    sd = cast_state_dict(sd, dtype)
    out = lstm_stack(sd, "lstm", x.to(dtype), num_layers=3, impl=lstm_impl)   # :20
    e = out[:, -1, :] @ sd["embedding.weight"].t() + sd["embedding.bias"]     # :21
    d_vec = e / e.norm(p=2, dim=-1, keepdim=True)                              # :22-23
    return e @ sd["output.weight"].t() + sd["output.bias"], d_vec             # :24-25
