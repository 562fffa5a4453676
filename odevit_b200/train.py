"""The reference's classification train loop (train.py:18-110) and a data-parallel launcher around it
(SURVEY section 8 row (f)4).

`train_classification_task` keeps the reference's signature and its step semantics -- verified against the
unmodified function on a CPU stand-in model (tests/test_train_loop.py):
  * the model is called as `model(**pixel_values, labels=labels, output_attentions=True,
    output_attention_trajectory=True)` (train.py:47-53);
  * the objective is `output["loss"] + 2 * jasmin_loss` when the model returns a JaSMin term (the loop adds it twice:
    train.py:60-66), accumulated IN PLACE on `output["loss"]`;
  * gradients accumulate over `num_accumulation_steps` batches without rescaling; then clip, `optimizer.step()`,
    `zero_grad()`, `scheduler.step()`; a partial group at the end of the epoch is left in `.grad` (train.py:80-87);
  * `params = model.parameters()` is a GENERATOR (train.py:32): `clip_grad_norm_` consumes it at the first optimizer
    step of the call, every later clip of the same call sees no parameters and does nothing.  `faithful_clip=True`
    (default) reproduces that; `False` clips every time.
Additions (all optional, defaults = reference behaviour): `grad_sync` runs before the clip (data-parallel all-reduce,
odevit_b200.dp.FlatGradAllReduce), `finish_batch` completes a deferred GPU collation (odevit_b200.data.Collator),
`output_attention_trajectory=False` spares the export of every attention map of the solve (the loop never reads them
when the model returns `jasmin_loss`).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 \\
        -m odevit_b200.train --config experiment_vit_edo.yaml [--synthetic 4096] [--epochs 2] [key=value ...]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from collections import defaultdict
from typing import Callable, Optional

import torch

device = "cuda" if torch.cuda.is_available() else "cpu"


def train_classification_task(dataloader, model: torch.nn.Module, optimizer, criterion: Optional[torch.nn.Module], scheduler,
                              wandb_logger=None, epoch: int = 0, num_accumulation_steps: int = 16, log_every: int = 5_000,
                              *, grad_sync: Optional[Callable[[], None]] = None, finish_batch: Optional[Callable] = None,
                              faithful_clip: bool = True, output_attention_trajectory: bool = True,
                              jasmin_fallback: Optional[Callable] = None, progress: bool = False):
    """train.py:18-110.  Returns (model, mean epoch loss)."""
    metrics_epoch = defaultdict(float)
    metrics_iter = defaultdict(float)
    cumulative = 0
    params = model.parameters()            # a generator, as in the reference (see the module docstring)
    model.train()
    it = enumerate(dataloader)
    if progress:
        import tqdm
        it = tqdm.tqdm(it, desc="Training Procedure", leave=True, position=1, total=len(dataloader))
    for batch_idx, data in it:
        cumulative += 1
        if finish_batch is not None:
            data = finish_batch(data)
        pixel_values = data["pixel_values"].to(device) if hasattr(data["pixel_values"], "to") else {
            k: v.to(device) for k, v in data["pixel_values"].items()}
        labels = data["labels"].to(device)
        output = model(**pixel_values, labels=labels, output_attentions=True,
                       output_attention_trajectory=output_attention_trajectory)
        preds = output["logits"]
        soft_pred = preds.softmax(dim=-1).argmax(dim=-1)
        loss = output["loss"]
        jasmin_loss = output.get("jasmin_loss", None)
        if jasmin_loss is not None:
            loss += jasmin_loss
        else:
            if jasmin_fallback is None:
                raise KeyError("the model returned no 'jasmin_loss' (the reference falls back to models.utils.jasmin_loss "
                               "on the last exported map: pass it as jasmin_fallback)")
            jasmin_loss = jasmin_fallback(output["attention_trajectory"][-1:], k=10)
        loss += jasmin_loss
        loss.backward()

        metrics_epoch["epoch_loss"] += loss.item()
        metrics_iter["iteration_loss"] += loss.item()
        acc = (soft_pred == labels).float().mean(-1)
        metrics_iter["iteration_acc"] += acc
        metrics_epoch["epoch_acc"] += acc
        metrics_iter["jasmin_loss"] += jasmin_loss.item()
        metrics_epoch["jasmin_loss"] += jasmin_loss.item()

        if cumulative >= num_accumulation_steps:
            if grad_sync is not None:
                grad_sync()
            torch.nn.utils.clip_grad_norm_(params if faithful_clip else list(model.parameters()), 1.0)
            optimizer.step()
            optimizer.zero_grad()
            if scheduler:
                scheduler.step()
            cumulative = 0

        metrics_iter.update({"train/lr": optimizer.param_groups[0]["lr"]})
        if ((batch_idx + 1) % log_every) == 0:
            if wandb_logger:
                metrics_iter = {f"train/{key}": value / log_every for key, value in metrics_iter.items()}
                wandb_logger.log(metrics_iter)
                metrics_iter = defaultdict(float)

    loss_to_return = metrics_epoch["epoch_loss"] / len(dataloader)
    if wandb_logger:
        metrics_epoch = {f"train/{key}": value / len(dataloader) for key, value in metrics_epoch.items()}
        metrics_epoch.update({"train/epoch": epoch})
        wandb_logger.log(metrics_epoch)
    return model, loss_to_return


# ------------------------------------------------------------------------------------------------------------------
# launcher: main_classification_ode.py:51-205 without hydra / wandb / checkpoints (out of scope)
# ------------------------------------------------------------------------------------------------------------------
def load_config(path: str, overrides=()) -> dict:
    """The shipped experiment YAMLs (configs/classification/*.yaml) as plain dicts; `a.b.c=value` overrides in hydra's
    command-line form (values parsed as YAML scalars)."""
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f)
    for ov in overrides:
        key, _, val = ov.partition("=")
        node = cfg
        parts = key.split(".")
        for k in parts[:-1]:
            node = node.setdefault(k, {})
        node[parts[-1]] = yaml.safe_load(val)
    return cfg


def restart_schedule(optimizer, epochs: int, steps_per_epoch: int):
    """main_classification_ode.py:148-166: cosine with hard restarts, 10 % warm-up, epochs // 100 cycles."""
    from transformers.optimization import get_cosine_with_hard_restarts_schedule_with_warmup
    total = epochs * steps_per_epoch
    return get_cosine_with_hard_restarts_schedule_with_warmup(optimizer, num_warmup_steps=int(0.1 * total),
                                                              num_training_steps=total, num_cycles=epochs // 100)


class SyntheticImages(torch.utils.data.Dataset):
    """uint8 RGB images + labels with the (PIL image, int) item type of torchvision's CIFAR datasets."""

    def __init__(self, n: int, size: int, num_classes: int, seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        self.x = torch.randint(0, 256, (n, size, size, 3), dtype=torch.uint8, generator=g)
        self.y = torch.randint(0, num_classes, (n,), generator=g)

    def __len__(self):
        return self.x.shape[0]

    def __getitem__(self, i):
        return self.x[i].numpy(), int(self.y[i])


def main(argv=None):
    import torch.distributed as dist
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler

    import odevit_b200 as ob
    from odevit_b200.data import Collator, GpuImageProcessor
    from odevit_b200.dp import FlatGradAllReduce

    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--synthetic", type=int, default=0, help="train on N synthetic uint8 images instead of cfg.data.dataset")
    ap.add_argument("--epochs", type=int, default=0, help="run only this many epochs (schedule still sized by the config)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("overrides", nargs="*")
    a = ap.parse_args(argv)
    cfg = load_config(a.config, a.overrides)
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    inputs = dict(cfg["modeling"]["inputs"])
    torch.manual_seed(0)                                             # replicas start from identical weights
    kind = cfg["modeling"].get("type", "vit")
    model = (ob.ViTNeuralODE if kind == "vit" else ob.macaron.ViTMacaron)(**inputs).cuda()
    model.precision = a.precision
    ds_cfg = cfg["data"]["dataset"]
    if a.synthetic:
        raw = 32 if "cifar" in str(ds_cfg.get("name", "")) else int(inputs["img_size"])
        train_dataset = SyntheticImages(a.synthetic, raw, int(inputs["num_classes"]), seed=1)
    elif ds_cfg["name"] in ("cifar100", "cifar10"):
        from torchvision.datasets import CIFAR10, CIFAR100
        train_dataset = (CIFAR100 if ds_cfg["name"] == "cifar100" else CIFAR10)(root=ds_cfg["dataset_path"], download=False, train=True)
    else:
        from torchvision.datasets import ImageFolder
        train_dataset = ImageFolder(root=ds_cfg["dataset_path"] + "/train")
    # facebook/dino-vitb16 preprocessor_config.json: size 224, resample 2 (bilinear), ImageNet mean / std
    processor = GpuImageProcessor(size=int(inputs["img_size"]), image_mean=(0.485, 0.456, 0.406), image_std=(0.229, 0.224, 0.225))
    collator = Collator(processor, defer=True)
    lc = dict(cfg["data"]["collator"]["train"])
    sampler = DistributedSampler(train_dataset, num_replicas=world, rank=rank, shuffle=bool(lc.pop("shuffle", True)),
                                 drop_last=bool(lc.get("drop_last", False))) if world > 1 else None
    loader = DataLoader(train_dataset, collate_fn=collator.classification_collate_fn, sampler=sampler,
                        shuffle=(sampler is None and bool(lc.pop("shuffle", True))), **{k: v for k, v in lc.items() if k != "shuffle"})
    optimizer = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=5e-2)          # main_classification_ode.py:141-146
    setup = cfg["setup"]["dict"]
    scheduler = restart_schedule(optimizer, int(setup["epochs"]), len(loader))
    sync = FlatGradAllReduce(model.parameters()) if world > 1 else None
    n_epochs = a.epochs or int(setup["epochs"]) - 1
    for epoch in range(1, n_epochs + 1):                              # :171-176 (epochs start at 1)
        if sampler is not None:
            sampler.set_epoch(epoch)
        t0 = time.time()
        _, loss = train_classification_task(loader, model, optimizer, None, scheduler, wandb_logger=None, epoch=epoch,
                                            num_accumulation_steps=int(setup["accumulation_steps"]),
                                            log_every=int(setup["log_every"]), grad_sync=sync, finish_batch=collator.finish,
                                            output_attention_trajectory=False)
        torch.cuda.synchronize()
        dt = time.time() - t0
        if rank == 0:
            n_img = len(loader) * int(lc.get("batch_size", 1)) * world
            print(json.dumps({"epoch": epoch, "loss": loss, "seconds": round(dt, 3), "img_per_s": round(n_img / dt, 1),
                              "world": world, "lr": optimizer.param_groups[0]["lr"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1:])
