from .losses import NTXentLoss, SupervisedContrastiveLoss, get_loss_fn
from .graph import GraphedTrainStep
from .optim import FusedClipAdam
from .trainer import ContrastiveTrainer

__all__ = ["NTXentLoss", "SupervisedContrastiveLoss", "get_loss_fn", "FusedClipAdam", "GraphedTrainStep", "ContrastiveTrainer"]
