"""Import-compatibility stub for the reference's network/layer.py, a half-ported TensorFlow helper file that no
module on the path imports.  Only `sigmoid_rampup` (layer.py:75-82, a duplicate of
trainer/baseTrainer.py:64-72) has a meaning on the path."""
import math


def sigmoid_rampup(current, rampup_length):
    if rampup_length == 0:
        return 1.0
    current = min(max(current, 0.0), rampup_length)
    phase = 1.0 - current / rampup_length
    return float(math.exp(-5.0 * phase * phase))
