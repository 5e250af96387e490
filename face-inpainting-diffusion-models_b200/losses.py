"""Output-parameterisation selectors used by the sampler dispatch.

Mirrors the three enums of the reference (`losses.py:10-39`); only the members and
`LossType.is_vb` are part of the sampling path.  The training-time KL / NLL helpers of
the reference (`losses.py:42-97`) are out of scope.
"""
from enum import Enum, auto


class ModelMeanType(Enum):
    PREVIOUS_X = auto()   # network predicts x_{t-1}
    START_X = auto()      # network predicts x_0
    EPSILON = auto()      # network predicts the noise


class ModelVarType(Enum):
    LEARNED = auto()
    FIXED_SMALL = auto()
    FIXED_LARGE = auto()
    LEARNED_RANGE = auto()


class LossType(Enum):
    MSE = auto()
    RESCALED_MSE = auto()
    KL = auto()
    RESCALED_KL = auto()

    def is_vb(self):
        return self in (LossType.KL, LossType.RESCALED_KL)
