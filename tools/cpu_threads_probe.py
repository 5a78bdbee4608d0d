import os, sys, time, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0,'/root/repo/tests')
import bench, argparse
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "cgroup", open("/sys/fs/cgroup/cpu.max").read().strip() if os.path.exists("/sys/fs/cgroup/cpu.max") else None, "threads", bench.host_threads())
args = argparse.Namespace(n_clips=1, n_frm=2, size=224, txt_len=32, n_ex=1)
for nt in (8, 16, 32, 64):
    os.environ["CB_CPU_THREADS"] = str(nt)
    sd, batch = bench._cpu_setup(args, 2)
    torch.set_num_threads(nt)
    bench._cpu_step(sd, batch, 1, 2, 224)
    t0=time.time(); bench._cpu_step(sd, batch, 1, 2, 224); print(nt, "threads: %.2f s/step (2 videos x 1 clip)" % (time.time()-t0), flush=True)
