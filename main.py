#!/usr/bin/env python
"""Drop-in for the reference's `main.py` (main.py:4-112): the same twelve command-line flags, the same construction order
(models -> optimizers -> checkpoint -> loaders -> ModelWrapper -> train / test), over the B200 package.

Differences, all at the places SURVEY.md section 8 replaces:
  * `--use_data_parallel` does not wrap the models in nn.DataParallel (main.py:91-94).  The script re-launches itself
    under torchrun with one process per GPU of `--gpus_to_use`; every rank builds the same models, takes its own shard of
    every batch (`distributed.shard_range`) and the gradients are averaged with NCCL (distributed.GradientReducer).
  * the optimizers are the fused multi-tensor Adam (same defaults and state_dict as torch.optim.Adam, checkpoints load
    both ways);
  * the data set comes from the user's own `data.py` (the reference's Places365 loader is outside this package's scope,
    SURVEY 8f); `--synthetic N` (an addition) trains on N seeded synthetic samples in the loader's format instead, which is
    what makes the script runnable without the data set;
  * `--test` calls `validate()` without the `device=` keyword the reference passes and its own `validate` does not accept
    (main.py:111 raises TypeError there), and needs a FID function (`--fid module:function`).
"""
import os
import subprocess
import sys
from argparse import ArgumentParser


def build_parser():
    parser = ArgumentParser()
    parser.add_argument('--train', default=False, action='store_true', help='Train network')
    parser.add_argument('--test', default=False, action='store_true', help='Test network')
    parser.add_argument('--batch_size', type=int, default=20,
                        help='Batch size of the training and test set, per GPU (default=20)')
    parser.add_argument('--lr', type=float, default=1e-05, help='Main learning rate of the adam optimizer (default=1e-05)')
    parser.add_argument('--channel_factor', type=float, default=1.0,
                        help='Channel factor adopts the number of channels utilized in G and D (default=1)')
    parser.add_argument('--device', type=str, default='cuda', help='Device to use (default=cuda)')
    parser.add_argument('--gpus_to_use', type=str, default='0', help='Indexes of the GPUs to be use (default=0)')
    parser.add_argument('--use_data_parallel', default=False, action='store_true',
                        help='Use multiple GPUs: one process per GPU of --gpus_to_use, NCCL gradient averaging')
    parser.add_argument('--load_checkpoint', type=str, default=None, help='Path to checkpoint to be loaded (default=None)')
    parser.add_argument('--load_pretrained_vgg16', type=str, default='pre_trained_models/vgg_places_365_fine_tuned.pt',
                        help='Name of the pretrained (places365) vgg16 network the be loaded from model file (.pt)')
    parser.add_argument('--path_to_places365', type=str, default='places365_standard', help='Path to places365 dataset.')
    parser.add_argument('--epochs', type=int, default=50, help='Epochs to perform while training (default=50)')
    # additions (not in the reference)
    parser.add_argument('--synthetic', type=int, default=0,
                        help='Train on this many seeded synthetic samples instead of Places365 (no data set needed)')
    parser.add_argument('--precision', type=str, default='bf16', choices=['bf16', 'split'],
                        help="bf16: throughput mode; split: hi+lo BF16 planes, matches the FP32 reference to ~1e-5")
    parser.add_argument('--fid', type=str, default=None, help='module:function computing the FID for --test / validation')
    return parser


def torchrun_command(argv, gpus):
    """The one-process-per-GPU launch that replaces nn.DataParallel (main.py:91-94)."""
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(len(gpus)),
            "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29533"),
            os.path.abspath(__file__)] + list(argv)


class SyntheticPlaces(object):
    """`len(dataset)` samples in the format of data.Places365.__getitem__ (data.py:39-65), seeded per index."""

    def __init__(self, length):
        self.length = length

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        import random
        import numpy as np
        import torch
        from semantic_pyramid_for_image_generation_b200 import misc
        g = torch.Generator().manual_seed(index)
        image = torch.rand(3, 256, 256, generator=g) * 2 - 1
        label = torch.nn.functional.one_hot(torch.randint(0, 365, (1,), generator=g)[0], 365).long()
        random.seed(index)
        np.random.seed(index)
        return image, label, misc.get_masks_for_training()


def collate(batch):
    """data.image_label_list_of_masks_collate_function (data.py:68-90) without the requires_grad flags nobody reads."""
    import torch
    images = torch.stack([b[0] for b in batch])
    labels = torch.stack([b[1] for b in batch])
    masks = [torch.stack([b[2][level] for b in batch]) for level in range(len(batch[0][2]))]
    return images, labels, masks


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    args = build_parser().parse_args(argv)
    gpus = [g for g in args.gpus_to_use.split(",") if g != ""]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.use_data_parallel and len(gpus) > 1 and world == 1:
        os.environ['CUDA_VISIBLE_DEVICES'] = args.gpus_to_use
        cmd = torchrun_command(argv, gpus)
        if os.environ.get("SPYR_MAIN_DRY_RUN"):
            print(" ".join(cmd))
            return 0
        return subprocess.call(cmd)
    if world == 1:
        os.environ['CUDA_VISIBLE_DEVICES'] = args.gpus_to_use

    import torch
    from torch.utils.data import DataLoader
    from semantic_pyramid_for_image_generation_b200 import distributed, ops
    from semantic_pyramid_for_image_generation_b200.model_wrapper import ModelWrapper
    from semantic_pyramid_for_image_generation_b200.models import Discriminator, Generator, VGG16
    from semantic_pyramid_for_image_generation_b200.optim import FusedAdam

    ops.set_precision(args.precision)
    reducer = distributed.init_from_env("nccl") if world > 1 else None
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    generator = Generator(channels_factor=args.channel_factor).cuda()
    discriminator = Discriminator(channel_factor=args.channel_factor).cuda()
    vgg16 = VGG16()
    if os.path.isfile(args.load_pretrained_vgg16):
        vgg16.load_state_dict(torch.load(args.load_pretrained_vgg16, map_location='cpu'))
    elif not args.synthetic:
        raise SystemExit("pretrained VGG-16 not found: %s" % args.load_pretrained_vgg16)
    generator_optimizer = FusedAdam(generator.parameters(), lr=args.lr)
    discriminator_optimizer = FusedAdam(discriminator.parameters(), lr=args.lr)
    if args.load_checkpoint is not None:
        checkpoint = torch.load(args.load_checkpoint, map_location='cpu')
        generator.load_state_dict(checkpoint['generator'])
        discriminator.load_state_dict(checkpoint['discriminator'])
        generator_optimizer.load_state_dict(checkpoint['generator_optimizer'])
        discriminator_optimizer.load_state_dict(checkpoint['discriminator_optimizer'])
    if reducer is None or reducer.rank == 0:
        print('Number of generator parameters', sum(p.numel() for p in generator.parameters()))
        print('Number of discriminator parameters', sum(p.numel() for p in discriminator.parameters()))

    if args.synthetic:
        rank = reducer.rank if reducer is not None else 0
        lo, hi = distributed.shard_range(args.synthetic, rank, world)
        train_set = torch.utils.data.Subset(SyntheticPlaces(args.synthetic), range(lo, hi))
        val_set = SyntheticPlaces(14)  # inference() draws 7 distinct validation samples (model_wrapper.py:258)
        collate_fn = collate
    else:
        import data  # the user's loader module (reference data.py); outside this package's scope
        train_set = data.Places365(path_to_index_file=args.path_to_places365, index_file_name='train.txt')
        val_set = data.Places365(path_to_index_file=args.path_to_places365, index_file_name='val.txt', max_length=6000,
                                 validation=True)
        collate_fn = data.image_label_list_of_masks_collate_function
        if world > 1:
            lo, hi = distributed.shard_range(len(train_set), reducer.rank, world)
            train_set = torch.utils.data.Subset(train_set, range(lo, hi))
    workers = min(args.batch_size, os.cpu_count() or 1)
    training_dataset = DataLoader(train_set, batch_size=args.batch_size, num_workers=workers, shuffle=True, drop_last=True,
                                  collate_fn=collate_fn, pin_memory=True)
    validation_dataset = DataLoader(val_set, batch_size=2 * args.batch_size, num_workers=workers, shuffle=True,
                                    collate_fn=collate_fn)
    fid_function = None
    if args.fid:
        import importlib
        module, function = args.fid.split(":")
        fid_function = getattr(importlib.import_module(module), function)
    model_wrapper = ModelWrapper(generator=generator, discriminator=discriminator, vgg16=vgg16,
                                 training_dataset=training_dataset, validation_dataset=validation_dataset,
                                 generator_optimizer=generator_optimizer, discriminator_optimizer=discriminator_optimizer,
                                 reducer=reducer, fid_function=fid_function)
    if args.train:
        model_wrapper.train(epochs=args.epochs, device=args.device)
    if args.test:
        print('FID=', model_wrapper.validate())
        model_wrapper.inference(device=args.device)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
