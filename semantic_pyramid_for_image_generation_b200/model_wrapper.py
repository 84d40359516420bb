"""Training / validation / inference driver, drop-in for the reference's `ModelWrapper` (model_wrapper.py:21-296).

Same constructor keywords and `train(epochs, validate_after_n_iterations, device, save_model_after_n_epochs, w_rec,
w_div)`, `validate()`, `inference(device)` entry points, same checkpoint dictionary and logged metric names.  The
iteration body (`training_step`) issues the reference's five forwards in the reference's order (G, D(real), D(fake), G,
D(fake)) but
  * keeps every tensor on the device: the five logged scalars leave the GPU in ONE copy per iteration;
  * does not compute what the reference computes and throws away (D weight gradients in the generator phase,
    image / mask gradients of the real batch);
  * replaces nn.DataParallel (main.py:91-94) by one process per GPU: pass `reducer=distributed.GradientReducer()`
    and gradients are averaged over ranks with NCCL after each backward.
"""
import os
from datetime import datetime
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import engine, misc
from .lossfunction import (DiversityLoss, LSGANDiscriminatorLoss, LSGANGeneratorLoss, SemanticReconstructionLoss,
                           lsgan_term)
from .models import VGG16

METRICS = ("loss_discriminator_real", "loss_discriminator_fake", "loss_generator",
           "loss_generator_semantic_reconstruction", "loss_generator_diversity")


def _unwrap(module):
    return module.module if isinstance(module, nn.DataParallel) else module


def _set_requires_grad(module, flag):
    for p in module.parameters():
        p.requires_grad_(flag)


class ModelWrapper(object):
    '''
    Model wrapper implementing training, validation and inference of the whole adversarial architecture
    '''

    def __init__(self, generator, discriminator, training_dataset, validation_dataset, vgg16=None,
                 generator_optimizer: torch.optim.Optimizer = None, discriminator_optimizer: torch.optim.Optimizer = None,
                 generator_loss: nn.Module = None, discriminator_loss: nn.Module = None,
                 semantic_reconstruction_loss: nn.Module = None, diversity_loss: nn.Module = None,
                 save_data_path: str = 'saved_data', reducer=None, fid_function=None) -> None:
        # nn.DataParallel wrappers are accepted for API compatibility and unwrapped: multi-GPU runs are one process per GPU
        self.generator = _unwrap(generator)
        self.discriminator = _unwrap(discriminator)
        self.training_dataset = training_dataset
        self.validation_dataset_fid = validation_dataset
        self.vgg16 = _unwrap(vgg16) if vgg16 is not None else VGG16()
        self.generator_optimizer = generator_optimizer
        self.discriminator_optimizer = discriminator_optimizer
        self.generator_loss = generator_loss or LSGANGeneratorLoss()
        self.discriminator_loss = discriminator_loss or LSGANDiscriminatorLoss()
        self.semantic_reconstruction_loss = semantic_reconstruction_loss or SemanticReconstructionLoss()
        self.diversity_loss = diversity_loss or DiversityLoss()
        self.latent_dimensions = self.generator.latent_dimensions
        self.reducer = reducer
        self.fid_function = fid_function
        _set_requires_grad(self.vgg16, False)  # frozen encoder (model_wrapper.py:67-68)
        self.logger = misc.Logger()
        self.is_main_process = reducer is None or reducer.rank == 0
        stamp = str(datetime.now())
        self.path_save_models = os.path.join(save_data_path, 'models_' + stamp)
        self.path_save_plots = os.path.join(save_data_path, 'plots_' + stamp)
        self.path_save_metrics = os.path.join(save_data_path, 'metrics_' + stamp)
        if self.is_main_process:
            for path in (self.path_save_models, self.path_save_plots, self.path_save_metrics):
                os.makedirs(path, exist_ok=True)
        for name in ('generator', 'discriminator', 'vgg16', 'generator_optimizer', 'discriminator_optimizer',
                     'generator_loss', 'discriminator_loss', 'diversity_loss', 'semantic_reconstruction_loss'):
            self.logger.hyperparameter[name] = str(getattr(self, name))
        self.progress_bar = None
        self._images_seen = 0

    # ------------------------------------------------------------------------------------------------
    # one iteration of model_wrapper.py:136-190
    # ------------------------------------------------------------------------------------------------
    def _second_stream(self, device, index=0):
        """Side streams for the network-level overlap inside a phase (None when SPYR_PHASE_STREAMS=0)."""
        if os.environ.get("SPYR_PHASE_STREAMS", "1") == "0":
            return None
        key = (device.type, device.index, index)
        streams = self.__dict__.setdefault("_phase_streams", {})
        if key not in streams:
            streams[key] = torch.cuda.Stream(device=device)
        return streams[key]

    def _phase_discriminator(self, images_real, labels, masks, z_d):
        """VGG(real), G (no grad), D(real), D(fake), LSGAN loss, backward.  Leaves D's gradients in D's flat arena."""
        G, D, V = self.generator, self.discriminator, self.vgg16
        batch, device = images_real.shape[0], images_real.device
        G.zero_grad(set_to_none=True)
        D.zero_grad(set_to_none=True)
        # Three streams.  `side`: VGG(real) -> G, both without autograd.  `third`: D(real) forward, its LSGAN term and
        # its whole backward.  This stream: D(fake) forward / loss / backward once the fake images exist.  The small maps
        # at the start of G and at both ends of D's backward leave most SMs idle; the 256x256 layers of the other
        # networks fill them.  D(fake)'s spectral-norm iteration must follow D(real)'s (model_wrapper.py:150-153 calls
        # them in that order on the same u, v): this stream waits for the end of D(real)'s forward.
        main, side, third = torch.cuda.current_stream(), self._second_stream(device), self._second_stream(device, 1)
        if side is not None:
            side.wait_stream(main)
            prep = self._second_stream(device, 2)
            if prep is not None and hasattr(G, "_sn") and os.environ.get("SPYR_SN_PREFETCH", "1") != "0":
                # G's spectral-norm pass does not need the VGG features: next to VGG(real) instead of behind it
                prep.wait_stream(main)
                with torch.no_grad(), torch.cuda.stream(prep):
                    engine.prepare_generator(G)
        with torch.no_grad(), torch.cuda.stream(side if side is not None else main):
            features_real = V(images_real)
            if z_d is None:
                z_d = torch.randn((batch, self.latent_dimensions), dtype=torch.float32, device=device)
            try:
                images_fake = G(input=z_d, features=features_real, masks=masks, class_id=labels.float())
            finally:
                G.__dict__.pop("_sn_prepared", None)  # never leave a prepared state behind for a later forward
        if side is None or type(self.discriminator_loss) is not LSGANDiscriminatorLoss:  # subclasses may override forward
            # a user-supplied loss couples the two predictions: D(real) still overlaps VGG -> G, one joint backward
            prediction_real = D(images_real, labels)
            if side is not None:
                main.wait_stream(side)
            prediction_fake = D(images_fake, labels)
            loss_d_real, loss_d_fake = self.discriminator_loss(prediction_real, prediction_fake)
            (loss_d_real + loss_d_fake).backward()
            return features_real, loss_d_real.detach(), loss_d_fake.detach()
        third.wait_stream(main)
        with torch.cuda.stream(third):
            prediction_real = D(images_real, labels)
            real_forward_done = torch.cuda.Event()
            real_forward_done.record(third)
            loss_d_real = lsgan_term(prediction_real, 1.0)  # the two LSGAN terms are independent (lossfunction.py:156-164)
            loss_d_real.backward()
        main.wait_stream(side)
        main.wait_event(real_forward_done)
        prediction_fake = D(images_fake, labels)
        loss_d_fake = lsgan_term(prediction_fake, 0.0)
        loss_d_fake.backward()  # adds into the arena of the first pass once `third` has finished (models._DiscriminatorFn)
        main.wait_stream(third)
        return features_real, loss_d_real.detach(), loss_d_fake.detach()

    def _generator_forward(self, features_real, labels, masks, z_g):
        """The part of the generator phase that does not touch D: under data parallelism it runs while D's gradients are
        still being all-reduced."""
        G = self.generator
        batch, device = labels.shape[0], labels.device
        G.zero_grad(set_to_none=True)
        if z_g is None:
            z_g = torch.randn((batch, self.latent_dimensions), dtype=torch.float32, device=device)
        return G(input=z_g, features=features_real, masks=masks, class_id=labels.float()), z_g

    def _phase_generator(self, features_real, labels, masks, z_g, w_rec, w_div, images_fake=None):
        """D's Adam step, then G, D(fake), the three generator losses, backward.  D acts as a fixed critic here, so its
        weight gradients are not requested (the reference computes and discards them).  With `images_fake`
        the generator forward has already run (`_generator_forward`, data-parallel schedule)."""
        G, D, V = self.generator, self.discriminator, self.vgg16
        device = labels.device
        main, side = torch.cuda.current_stream(), self._second_stream(device)
        overlapped = side if side is not None else main
        deferred_d_step = images_fake is not None
        if images_fake is None:
            # D's Adam step (HBM-bound, reads D's gradient arena) only has to finish before D is used again: it shares the
            # GPU with the generator forward, whose first layers are 4x4 ... 32x32 maps
            if side is not None:
                side.wait_stream(main)
            with torch.cuda.stream(overlapped):
                self.discriminator_optimizer.step()
            D.zero_grad(set_to_none=True)
            images_fake, z_g = self._generator_forward(features_real, labels, masks, z_g)
            if side is not None:
                main.wait_stream(side)  # D's new weights
        # VGG(fake) on the second stream next to D(fake); autograd runs each backward on its forward's stream, so the two
        # input-gradient chains overlap as well and meet at images_fake
        if side is not None:
            side.wait_stream(main)
        with torch.cuda.stream(overlapped):
            features_fake = V(images_fake)
        if deferred_d_step:
            self.discriminator_optimizer.step()  # data-parallel schedule: after the all-reduce, next to VGG(fake)
            D.zero_grad(set_to_none=True)
        _set_requires_grad(D, False)
        try:
            prediction_fake = D(images_fake, labels)
            loss_g = self.generator_loss(prediction_fake)
            loss_div = w_div * self.diversity_loss(images_fake, z_g)
            if side is not None:
                main.wait_stream(side)
            loss_rec = w_rec * self.semantic_reconstruction_loss(features_real, features_fake, masks)
            (loss_g + loss_rec + loss_div).backward()
        finally:
            _set_requires_grad(D, True)
        return loss_g.detach(), loss_rec.detach().reshape(()), loss_div.detach()

    def training_step(self, images_real: torch.Tensor, labels: torch.Tensor, masks: List[torch.Tensor],
                      w_rec: float = 0.1, w_div: float = 0.1, noise=None) -> Dict[str, torch.Tensor]:
        """Runs the discriminator update then the generator update on one device-resident batch and returns the five
        loss scalars as device tensors (no host synchronisation).  `noise` optionally supplies the two latent batches."""
        z_d, z_g = noise if noise is not None else (None, None)
        if self.reducer is not None and self.reducer.active and not self.__dict__.get("_replicas_synced", False):
            # rank 0's weights and buffers to every rank once (nn.DataParallel re-broadcast them every forward,
            # main.py:91-94): the replicas start identical even if the processes were seeded differently
            for module in (self.generator, self.discriminator, self.vgg16):
                self.reducer.broadcast_module(module)
            self._replicas_synced = True
        features_real, loss_d_real, loss_d_fake = self._phase_discriminator(images_real, labels, masks, z_d)
        if self.reducer is not None and self.reducer.active:
            # D's gradient all-reduce runs on the collective stream while the generator forward (which does not use D)
            # runs here; D's Adam step follows the all-reduce
            self.reducer.average(self.discriminator, wait=False)
            images_fake, z_g = self._generator_forward(features_real, labels, masks, z_g)
            self.reducer.wait()
            loss_g, loss_rec, loss_div = self._phase_generator(features_real, labels, masks, z_g, w_rec, w_div,
                                                               images_fake=images_fake)
        else:
            loss_g, loss_rec, loss_div = self._phase_generator(features_real, labels, masks, z_g, w_rec, w_div)
        if self.reducer is not None:
            self.reducer.average(self.generator)
        self.generator_optimizer.step()
        return {"loss_discriminator_real": loss_d_real, "loss_discriminator_fake": loss_d_fake, "loss_generator": loss_g,
                "loss_generator_semantic_reconstruction": loss_rec, "loss_generator_diversity": loss_div}

    def capture_training_step(self, images_real, labels, masks, w_rec: float = 0.1, w_div: float = 0.1):
        """Returns a `CapturedTrainingStep`: the iteration recorded as CUDA graphs over the given (static) input tensors."""
        return CapturedTrainingStep(self, images_real, labels, masks, w_rec, w_div)

    def _to_device(self, images_real, labels, masks, device):
        images_real = images_real.detach().to(device, non_blocking=True)
        labels = labels.to(device, non_blocking=True)
        masks = [m.detach().to(device, non_blocking=True) for m in masks]
        return images_real, labels, masks

    def train(self, epochs: int = 20, validate_after_n_iterations: int = 100000, device: str = 'cuda',
              save_model_after_n_epochs: int = 1, w_rec: float = 0.1, w_div: float = 0.1) -> None:
        """
        Training loop (reference model_wrapper.py:93-228)
        """
        self.logger.hyperparameter['w_rec'] = str(w_rec)
        self.logger.hyperparameter['w_div'] = str(w_div)
        batch_size = self.training_dataset.batch_size
        validate_after_n_iterations = max(1, validate_after_n_iterations // batch_size) * batch_size
        self.generator.train()
        self.discriminator.train()
        self.vgg16.eval()
        for module in (self.generator, self.discriminator, self.vgg16):
            module.to(device)
        total = epochs * len(self.training_dataset.dataset)
        try:
            from tqdm import tqdm
            self.progress_bar = tqdm(total=total, dynamic_ncols=True, disable=not self.is_main_process)
        except Exception:  # pragma: no cover
            self.progress_bar = None
        fid = float('nan')
        if self.validation_dataset_fid is not None and self.fid_function is not None:
            self.inference(device=device)
            fid = self.validate()
        for epoch in range(epochs):
            self.generator.train()
            self.discriminator.train()
            self.vgg16.eval()
            for images_real, labels, masks in self.training_dataset:
                self._images_seen += images_real.shape[0]
                if self.progress_bar is not None:
                    self.progress_bar.update(n=images_real.shape[0])
                images_real, labels, masks = self._to_device(images_real, labels, masks, device)
                losses = self.training_step(images_real, labels, masks, w_rec=w_rec, w_div=w_div)
                values = torch.stack([losses[name] for name in METRICS]).tolist()  # the only host sync of the iteration
                if self.progress_bar is not None:
                    self.progress_bar.set_description(
                        'FID={:.4f}, Loss Div={:.4f}, Loss Rec={:.4f}, Loss G={:.4f}, Loss D={:.4f}'.format(
                            fid, values[4], values[3], values[2], values[0] + values[1]))
                for name, value in zip(METRICS, values):
                    self.logger.log(metric_name=name, value=value)
                self.logger.log(metric_name='iterations', value=self._images_seen)
                self.logger.log(metric_name='epoch', value=epoch)
                if self._images_seen % validate_after_n_iterations == 0 and self.fid_function is not None \
                        and self.validation_dataset_fid is not None:
                    fid = self.validate()
                    self.inference(device=device)
                    self.logger.log(metric_name='fid', value=fid)
                    self.logger.log(metric_name='iterations_fid', value=self._images_seen)
                    if self.is_main_process:
                        self.logger.save_metrics(self.path_save_metrics)
            if epoch % save_model_after_n_epochs == 0 and self.is_main_process:
                self.save_checkpoint(os.path.join(self.path_save_models, 'checkpoint_{}.pt'.format(str(epoch).zfill(3))))
            if self.validation_dataset_fid is not None:
                self.inference(device=device)
            if self.is_main_process:
                self.logger.save_metrics(self.path_save_metrics)
        if self.progress_bar is not None:
            self.progress_bar.close()

    def save_checkpoint(self, path: str) -> None:
        """Same dictionary as reference model_wrapper.py:215-223 (state_dict keys are interchangeable)."""
        torch.save({"generator": self.generator.state_dict(), "discriminator": self.discriminator.state_dict(),
                    "generator_optimizer": self.generator_optimizer.state_dict(),
                    "discriminator_optimizer": self.discriminator_optimizer.state_dict()}, path)

    @torch.no_grad()
    def validate(self) -> float:
        '''
        FID estimate through the user-supplied `fid_function(dataset_real=, generator=, vgg16=)` (the reference's
        frechet_inception_distance needs a pretrained InceptionV3 download and is outside this package's scope)
        '''
        if self.fid_function is None:
            raise RuntimeError("validate() needs ModelWrapper(..., fid_function=...): the FID of the reference depends "
                               "on downloaded InceptionV3 weights (frechet_inception_distance.py:12-42)")
        self.generator.eval()
        self.vgg16.eval()
        fid = self.fid_function(dataset_real=self.validation_dataset_fid, generator=self.generator, vgg16=self.vgg16)
        self.generator.train()
        return float(fid)

    @torch.no_grad()
    def inference(self, device: str = 'cuda') -> Optional[torch.Tensor]:
        '''
        7x7 grid: seven validation images (rows) generated from each of the seven pyramid levels (columns), as
        reference model_wrapper.py:247-296 but one batched generator call per level instead of 49 single-image calls
        '''
        import numpy as np
        self.generator.to(device)
        self.vgg16.to(device)
        self.generator.eval()
        dataset = self.validation_dataset_fid.dataset
        picks = np.random.choice(range(len(dataset)), replace=False, size=7)
        samples = [dataset[int(i)] for i in picks]
        images = torch.stack([s[0] for s in samples]).float().to(device)
        labels = torch.stack([s[1] for s in samples]).to(device)
        features = self.vgg16(images)
        grid = torch.empty(7, 7, images.shape[1], images.shape[2], images.shape[3], dtype=torch.float32, device=device)
        for level in range(7):
            masks = [m.expand(7, *m.shape[1:]).contiguous()
                     for m in misc.get_masks_for_inference(level, add_batch_size=True, device=device)]
            z = torch.randn(7, self.latent_dimensions, dtype=torch.float32, device=device)
            grid[:, level] = self.generator(input=z, features=features, masks=masks, class_id=labels.float())
        fake_images = grid.reshape(49, *grid.shape[2:])
        if self.is_main_process:
            try:
                import torchvision
                torchvision.utils.save_image(misc.normalize_0_1_batch(fake_images), os.path.join(
                    self.path_save_plots, 'predictions_{}.png'.format(self._images_seen)), nrow=7)
            except Exception:  # pragma: no cover - plotting is best effort
                pass
        self.generator.train()
        return fake_images


def red_active(wrapper):
    return wrapper.reducer is not None and wrapper.reducer.active


class CapturedTrainingStep(object):
    """One training iteration as three CUDA graphs with the two gradient all-reduces between them:

        graph A  discriminator phase (forward x4, backward)          -> NCCL average of D's gradient arena
        graph B  D Adam step + generator phase (forward x3, backward) -> NCCL average of G's gradient arena
        graph C  G Adam step

    All kernels of the step are launched through the C-ABI on the capturing stream and nothing synchronises, so the
    ~700 launches of an iteration cost three graph launches.  Inputs are read from the tensors given at capture time:
    copy new batches into them (`load`) and call the object.  Single-process runs skip the collectives."""

    def __init__(self, wrapper, images_real, labels, masks, w_rec=0.1, w_div=0.1):
        self.w = wrapper
        self.images, self.labels, self.masks = images_real, labels, list(masks)
        self._stage, self._pending = None, False
        # warm-up on a side stream (allocator, lazy tables, optimizer state), as CUDA-graph capture requires
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                wrapper.training_step(self.images, self.labels, self.masks, w_rec=w_rec, w_div=w_div)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _native
        self.graph_a, self.graph_b, self.graph_c = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        pool = torch.cuda.graph_pool_handle()
        _native.launch_count_reset()
        with torch.cuda.graph(self.graph_a, pool=pool):
            self.features_real, l_real, l_fake = wrapper._phase_discriminator(self.images, self.labels, self.masks, None)
        self.d_arena = wrapper.discriminator._last_grad_arena
        if red_active(wrapper) and not wrapper.reducer.grads_alias_arena(wrapper.discriminator, self.d_arena):
            raise RuntimeError("captured step: the discriminator's .grad tensors are not views of its gradient arena")
        self.graph_b1 = None
        red = wrapper.reducer
        if red is not None and red.active:
            # data-parallel schedule: graph B1 (generator forward) overlaps the all-reduce of D's gradients
            self.graph_b1 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_b1, pool=pool):
                images_fake, z_g = wrapper._generator_forward(self.features_real, self.labels, self.masks, None)
            with torch.cuda.graph(self.graph_b, pool=pool):
                l_g, l_rec, l_div = wrapper._phase_generator(self.features_real, self.labels, self.masks, z_g, w_rec, w_div,
                                                             images_fake=images_fake)
        else:
            with torch.cuda.graph(self.graph_b, pool=pool):
                l_g, l_rec, l_div = wrapper._phase_generator(self.features_real, self.labels, self.masks, None, w_rec,
                                                             w_div)
        self.g_arena = wrapper.generator._last_grad_arena
        if red_active(wrapper) and not wrapper.reducer.grads_alias_arena(wrapper.generator, self.g_arena):
            raise RuntimeError("captured step: the generator's .grad tensors are not views of its gradient arena")
        with torch.cuda.graph(self.graph_c, pool=pool):
            wrapper.generator_optimizer.step()
        self.launches_per_step = _native.launch_count()
        self.losses = {"loss_discriminator_real": l_real, "loss_discriminator_fake": l_fake, "loss_generator": l_g,
                       "loss_generator_semantic_reconstruction": l_rec, "loss_generator_diversity": l_div}

    def load(self, images_real, labels, masks) -> None:
        """Asynchronous copy of a new batch (pinned host or device tensors) into the captured input buffers."""
        self.images.copy_(images_real, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        for dst, src in zip(self.masks, masks):
            dst.copy_(src, non_blocking=True)

    def prefetch(self, images_real, labels, masks) -> None:
        """Starts the host-to-device copy of the NEXT batch on a copy stream into staging buffers, so that it crosses
        PCIe while the current iteration computes; the next call moves it into the captured inputs device-to-device
        (17.9 MB at bs 20: ~10 us instead of ~0.4 ms of PCIe time on the critical path)."""
        if self._stage is None:
            self._stage = (torch.empty_like(self.images), torch.empty_like(self.labels),
                           [torch.empty_like(m) for m in self.masks])
            self._copy_stream = torch.cuda.Stream(device=self.images.device)
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
        cs = self._copy_stream
        cs.wait_event(self._consumed)  # the previous batch has left the staging buffers
        with torch.cuda.stream(cs):
            self._stage[0].copy_(images_real, non_blocking=True)
            self._stage[1].copy_(labels, non_blocking=True)
            for dst, src in zip(self._stage[2], masks):
                dst.copy_(src, non_blocking=True)
            self._staged.record(cs)
        self._pending = True

    def __call__(self) -> Dict[str, torch.Tensor]:
        if self._pending:
            main = torch.cuda.current_stream()
            main.wait_event(self._staged)
            self.load(*self._stage)
            self._consumed.record(main)
            self._pending = False
        red = self.w.reducer
        self.graph_a.replay()
        if self.graph_b1 is not None:
            red.average_flat(self.d_arena, wait=False)
            self.graph_b1.replay()
            red.wait()
        self.graph_b.replay()
        if red is not None and red.active:
            red.average_flat(self.g_arena)
        self.graph_c.replay()
        return self.losses
