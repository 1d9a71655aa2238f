"""Checkpoint loading for the drop-in scripts (train_resnet.py:215-228, decode.py:150-160 in the reference).

The reference's checkpoints are `{epoch, arch, state_dict, best_acc1, optimizer}` — tensors, scalars and strings only —
so they load under torch's restricted unpickler (`weights_only=True`), which cannot execute code embedded in an untrusted
model file.  A legacy file that needs the full unpickler is refused unless SVK_UNSAFE_LOAD=1 is set explicitly."""
import os
import pickle

import torch


def load_checkpoint(path, map_location=None):
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except pickle.UnpicklingError as ex:
        if os.environ.get("SVK_UNSAFE_LOAD", "0") != "1":
            raise RuntimeError("%s does not load with weights_only=True (%s); set SVK_UNSAFE_LOAD=1 to unpickle it anyway "
                               "if you trust the file" % (path, str(ex).splitlines()[0]))
        return torch.load(path, map_location=map_location, weights_only=False)
