"""load_checkpoint_flexible (reference: models/__init__.py:20-51): load third-party checkpoints whose
keys may or may not carry DDP's ``module.`` prefix; unmatched keys are ignored, missing keys keep
their current values."""
from __future__ import annotations

import torch


def load_checkpoint_flexible(model, checkpoint_path, state_dict_key=None):
    blob = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
    state = blob[state_dict_key] if state_dict_key is not None else blob
    own = model.state_dict()
    for key, value in state.items():
        for cand in (key, key[7:] if key.startswith("module.") else "module." + key):
            if cand in own:
                own[cand] = value
                break
    missing, unexpected = model.load_state_dict(own)
    if missing:
        print("Missing keys: ", ",".join(missing))
    if unexpected:
        print("Unexpected keys: ", ",".join(unexpected))
    return model
