"""Static SASS instruction count per source function / line (nvdisasm -g of the in-tree library)."""
import collections, os, re, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "mujoco_rl_environment_wrapper_b200", "libmjb.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = os.path.join(tmp, "mjb_batch.sm_100a.cubin")
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
cur, cnt = None, collections.Counter()
for l in dis.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l) and cur:
        cnt[cur] += 1
print("static instructions:", sum(cnt.values()))
fn_starts = []
for fname in ('step_kernel.cuh', 'env_kernel.cuh', 'warp_prims.cuh', 'mjb_batch.cu'):
    for i, l in enumerate(open(os.path.join(root, 'mujoco_rl_environment_wrapper_b200', 'csrc', fname)), 1):
        m = re.match(r'^(?:MJB_DEV_NOINLINE|MJB_DEV|MJB_HD|__global__|__device__ __forceinline__)\s+[\w:<>\*& ]+?\s+\**(\w+)\s*\(', l)
        if m:
            fn_starts.append((fname, i, m.group(1)))
def fn_of(f, ln):
    best = None
    for (ff, i, n) in fn_starts:
        if ff == f and i <= ln:
            best = n
    return best
agg = collections.Counter()
for (f, ln), v in cnt.items():
    agg[(f, fn_of(f, ln))] += v
for k, v in agg.most_common(int(sys.argv[1]) if len(sys.argv) > 1 else 20):
    print(k, v)
print("top lines")
for k, v in cnt.most_common(15):
    print(k, v)
