"""How many select rounds does the decode chain need with the previous-cutoff hint?  (CTA 0 of the first problem of every
fused launch stores its round count; read after each token = the last layer's w2.)"""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from effort_b200 import ops  # noqa: E402
from effort_b200.model import DecodeModel, MistralConfig  # noqa: E402

m = DecodeModel.random_init(MistralConfig(n_layers=4, vocab=4096, max_seq=128), seed=3)
m.set_graphs(False)
ctx = ops.default_context()
for effort in (0.25, 0.5):
    m.reset()
    rounds = []
    tok = torch.tensor([1], dtype=torch.int32, device="cuda")
    for t in range(24):
        m.step(tok if t == 0 else None, effort)
        loops = C.c_int(0)
        ctx._L.effort_read_dispatch(ctx._h, None, 0, None, None, None, C.byref(loops), ops._stream_ptr())
        rounds.append(loops.value)
    print("effort", effort, "select rounds of the last fused launch per token:", rounds)
