mkdir -p gpurun_out
{
echo "== cgroup"; cat /proc/self/cgroup; for f in /sys/fs/cgroup/cpu.max /sys/fs/cgroup/cpu.stat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpu.weight /sys/fs/cgroup/memory.max; do echo "-- $f"; cat $f 2>&1; done
echo "== nproc $(nproc) ; sched_getaffinity"; python -c "import os; print(len(os.sched_getaffinity(0)), os.cpu_count())"
echo "== lscpu"; lscpu | head -25
echo "== cpu.stat before"; cat /sys/fs/cgroup/cpu.stat 2>&1
BCI_ZYGOTE_POOL=0 timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 4 --requests 2048 --scripts pass | cut -c1-300
echo "== cpu.stat after"; cat /sys/fs/cgroup/cpu.stat 2>&1
echo "== spin test: 64 busy processes for 2 s, how much CPU do they get"
python - <<'PY'
import multiprocessing as mp, time, os
def burn(q):
    t0=time.process_time(); w0=time.time()
    while time.time()-w0 < 2.0: pass
    q.put(time.process_time()-t0)
for n in (8, 32, 64, 128):
    q=mp.Queue(); ps=[mp.Process(target=burn,args=(q,)) for _ in range(n)]
    [p.start() for p in ps]; r=[q.get() for _ in ps]; [p.join() for p in ps]
    print(n, "procs: cpu seconds per 2 s wall each:", round(sum(r)/n,3), "total cores busy:", round(sum(r)/2.0,1))
PY
} > gpurun_out/r2i_cgroup.log 2>&1
cat gpurun_out/r2i_cgroup.log
