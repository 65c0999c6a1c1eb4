"""Self-generated regression fixture for the STDiT3 oracle (NOT a reference vector: STDiT3 is absent from
/root/reference, so this only guards the restatement against accidental edits — parity unpinned)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import stdit3_oracle as O  # noqa: E402

cfg = O.STDiT3_XS_2_config()
m = O.STDiT3(cfg).eval()
O.init_synthetic_weights(m)
with torch.no_grad():
    out = m(**O.synthetic_inputs(cfg, 1, 8, 16, 16))
np.savez_compressed(os.path.join(HERE, "stdit3_xs_selfcheck.npz"), out=out.numpy()[:, :, ::2, ::4, ::4])
print("ok", out.shape)
