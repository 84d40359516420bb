"""Fine-tuning of the VGG-16 pyramid encoder as a Places365 classifier on the B200 kernels.

Same procedure as the reference's vgg_16_train.py: `VGG16(return_output=True)` trained with cross entropy and Adam
(vgg_16_train.py:104-106), learning rate divided by ten every 30 epochs (:252-256), precision@1 / @5 bookkeeping
(:259-272), latest / best checkpoints (:227-230).  What runs on the device is this package's VGG16 -- tensor-core
convolutions forward, input- and weight-gradient, dropout in the classifier -- and `optim.FusedAdam`; the loss on the
(B, 365) logits is torch's own `CrossEntropyLoss`, as in the reference.  There is no CPU fallback.
"""
import shutil
import time
from typing import Iterable, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .models import VGG16
from .optim import FusedAdam


class RunningMean(object):
    """Last value and running average of a logged quantity (what the reference's AverageMeter exposes: val / avg)."""

    def __init__(self) -> None:
        self.val, self.sum, self.count = 0.0, 0.0, 0

    @property
    def avg(self) -> float:
        return self.sum / self.count if self.count else 0.0

    def update(self, value: float, n: int = 1) -> None:
        self.val = value
        self.sum += value * n
        self.count += n


def precision_at_k(logits: torch.Tensor, target: torch.Tensor, topk: Sequence[int] = (1,)) -> Tuple[torch.Tensor, ...]:
    """Percentage of samples whose label is among the k largest logits, for every k (vgg_16_train.py:259-272)."""
    ranked = logits.topk(max(topk), dim=1).indices
    hit = ranked.eq(target.view(-1, 1))
    return tuple(hit[:, :k].any(dim=1).float().mean() * 100.0 for k in topk)


def learning_rate_for_epoch(base_lr: float, epoch: int) -> float:
    return base_lr * (0.1 ** (epoch // 30))


def set_learning_rate(optimizer: torch.optim.Optimizer, lr: float) -> None:
    for group in optimizer.param_groups:
        group["lr"] = lr


def build(path_to_pre_trained_model: Optional[str] = None, lr: float = 1e-4, device: str = "cuda"):
    """Model, criterion and optimizer as main() of the reference constructs them (vgg_16_train.py:59,104-106)."""
    model = VGG16(path_to_pre_trained_model, return_output=True).to(device)
    criterion = nn.CrossEntropyLoss().to(device)
    optimizer = FusedAdam(model.parameters(), lr=lr)
    return model, criterion, optimizer


def train_epoch(loader: Iterable, model: VGG16, criterion: nn.Module, optimizer: torch.optim.Optimizer, epoch: int,
                print_freq: int = 10, device: str = "cuda", log=print):
    """One epoch of vgg_16_train.py:134-180.  Returns the (loss, prec@1, prec@5) running means."""
    losses, top1, top5, batch_time = RunningMean(), RunningMean(), RunningMean(), RunningMean()
    model.train()
    end = time.time()
    for i, (images, target) in enumerate(loader):
        images = images.to(device, non_blocking=True)
        target = target.to(device, non_blocking=True)
        logits = model(images)
        loss = criterion(logits, target)
        p1, p5 = precision_at_k(logits.detach(), target, (1, 5))
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
        n = images.shape[0]
        losses.update(float(loss.detach()), n)
        top1.update(float(p1), n)
        top5.update(float(p5), n)
        batch_time.update(time.time() - end)
        end = time.time()
        if log is not None and i % print_freq == 0:
            log("Epoch: [%d][%d]\tTime %.3f (%.3f)\tLoss %.4f (%.4f)\tPrec@1 %.3f (%.3f)\tPrec@5 %.3f (%.3f)" %
                (epoch, i, batch_time.val, batch_time.avg, losses.val, losses.avg, top1.val, top1.avg, top5.val, top5.avg))
    return losses, top1, top5


def validate(loader: Iterable, model: VGG16, criterion: nn.Module, print_freq: int = 10, device: str = "cuda",
             log=print) -> float:
    """vgg_16_train.py:183-224: eval mode, no gradients; returns the average precision@1."""
    losses, top1, top5 = RunningMean(), RunningMean(), RunningMean()
    model.eval()
    with torch.no_grad():
        for i, (images, target) in enumerate(loader):
            images = images.to(device, non_blocking=True)
            target = target.to(device, non_blocking=True)
            logits = model(images)
            loss = criterion(logits, target)
            p1, p5 = precision_at_k(logits, target, (1, 5))
            n = images.shape[0]
            losses.update(float(loss), n)
            top1.update(float(p1), n)
            top5.update(float(p5), n)
            if log is not None and i % print_freq == 0:
                log("Test: [%d]\tLoss %.4f (%.4f)\tPrec@1 %.3f (%.3f)\tPrec@5 %.3f (%.3f)" %
                    (i, losses.val, losses.avg, top1.val, top1.avg, top5.val, top5.avg))
    if log is not None:
        log(" * Prec@1 %.3f Prec@5 %.3f" % (top1.avg, top5.avg))
    return top1.avg


def save_checkpoint(state: dict, is_best: bool, filename: str = "checkpoint.pth.tar") -> None:
    torch.save(state, filename + "_latest.pth.tar")
    if is_best:
        shutil.copyfile(filename + "_latest.pth.tar", filename + "_best.pth.tar")


def fit(train_loader: Iterable, val_loader: Optional[Iterable], model: VGG16, criterion: nn.Module,
        optimizer: torch.optim.Optimizer, epochs: int = 3, start_epoch: int = 0, base_lr: float = 1e-4,
        checkpoint_name: Optional[str] = "VGG_16", device: str = "cuda", log=print) -> float:
    """The epoch loop of vgg_16_train.py:112-131; returns the best precision@1 seen."""
    best = 0.0
    if val_loader is not None:
        validate(val_loader, model, criterion, device=device, log=log)
    for epoch in range(start_epoch, epochs):
        set_learning_rate(optimizer, learning_rate_for_epoch(base_lr, epoch))
        train_epoch(train_loader, model, criterion, optimizer, epoch, device=device, log=log)
        prec1 = validate(val_loader, model, criterion, device=device, log=log) if val_loader is not None else 0.0
        is_best = prec1 > best
        best = max(best, prec1)
        if checkpoint_name is not None:
            save_checkpoint({"epoch": epoch + 1, "state_dict": model.state_dict(), "best_prec1": best}, is_best,
                            checkpoint_name)
    return best
