"""Developer tool: GPU timeline of one captured training step (torch.profiler / CUPTI): every kernel with its start, duration,
stream and grid, written as JSON lines for offline analysis of what overlaps with what (tools/timeline_report.py)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from semantic_pyramid_for_image_generation_b200 import models, ops
from semantic_pyramid_for_image_generation_b200.model_wrapper import ModelWrapper
from semantic_pyramid_for_image_generation_b200.optim import FusedAdam

out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.jsonl"
from semantic_pyramid_for_image_generation_b200 import distributed
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
reducer = distributed.init_from_env("nccl") if world > 1 else None
torch.manual_seed(0)
dev = torch.device("cuda", local)
G = models.Generator(channels_factor=1).to(dev).train()
D = models.Discriminator(channel_factor=1).to(dev).train()
V = models.VGG16().to(dev).eval()
w = ModelWrapper(G, D, None, None, vgg16=V, generator_optimizer=FusedAdam(G.parameters(), lr=1e-5),
                 discriminator_optimizer=FusedAdam(D.parameters(), lr=1e-5), save_data_path="/tmp/spyr_tl_%d" % rank,
                 reducer=reducer)
h_images, h_labels, h_masks = bench.host_batch(20, seed=rank)
imgs, labs, masks = h_images.to(dev), h_labels.to(dev), [m.to(dev) for m in h_masks]
for _ in range(3):
    w.training_step(imgs, labs, masks)
cap = w.capture_training_step(imgs, labs, masks)
for _ in range(5):
    cap()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        cap()
    torch.cuda.synchronize()
if rank != 0:
    torch.distributed.destroy_process_group()
    sys.exit(0)
prof.export_chrome_trace("/tmp/trace.json")
tr = json.load(open("/tmp/trace.json"))
n = 0
with open(out_path, "w") as f:
    for e in tr["traceEvents"]:
        if e.get("cat") == "kernel":
            a = e.get("args", {})
            f.write(json.dumps({"name": e["name"][:80], "ts": e["ts"], "dur": e["dur"], "stream": a.get("stream"),
                                "grid": a.get("grid"), "block": a.get("block"), "smem": a.get("shared memory")}) + "\n")
            n += 1
print("kernels", n)
if world > 1:
    torch.distributed.destroy_process_group()
