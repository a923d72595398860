"""2-GPU data-parallel smoke: eager step, graph capture, replays; prints progress (debug aid)."""
import faulthandler, os, sys, time
faulthandler.dump_traceback_later(45, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def P(*a):
    print(f"[r{rank} {time.time() % 1000:7.2f}]", *a, flush=True)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
P("pg ready")
from cloudaae_b200.train import CloudAAETrainer
from cloudaae_b200.synthesis import SegmentSynthesizer, load_models_xyz
B = 32
tr = CloudAAETrainer(batch_size=B, num_point=256, device=torch.device("cuda", lr), seed=rank, process_group=dist.group.WORLD)
P("trainer ready; params equal across ranks?")
chk = tr.v.flat.double().sum().reshape(1).clone(); lst = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(lst, chk); P([x.item() for x in lst])
syn = SegmentSynthesizer(load_models_xyz(device=torch.device("cuda", lr)), B, 256, seed=rank)
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ycb_poses.npz"))
sel = np.random.default_rng(rank).integers(0, len(z["class_id"]), B)
c = torch.from_numpy(z["class_id"][sel].astype(np.int32)).cuda(); a = torch.from_numpy(z["axisangle"][sel]).cuda(); t = torch.from_numpy(z["translation"][sel]).cuda()
l = tr.train_step_online(syn, c, a, t); torch.cuda.synchronize(); P("eager step ok", l.tolist())
chk = tr.v.flat.double().sum().reshape(1).clone(); dist.all_gather(lst, chk); P("params after step", [x.item() for x in lst])
tr.capture_online(syn, c, a, t); torch.cuda.synchronize(); P("capture ok")
for i in range(3):
    tr.replay(); torch.cuda.synchronize(); P("replay", i, tr.losses.tolist())
chk = tr.v.flat.double().sum().reshape(1).clone(); dist.all_gather(lst, chk); P("params after replays", [x.item() for x in lst])
dist.destroy_process_group(); P("done")
