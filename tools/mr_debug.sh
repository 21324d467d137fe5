#!/bin/bash
# manual 2-GPU probe: per-step statistics of the multirank worker in several modes
cd "$(dirname "$0")/.."
python tests/multirank_worker.py --out /tmp/n1 > /dev/null 2>&1
for mode in "peer 0" "peer 1" "nccl 0"; do
  set -- $mode
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29655 tests/multirank_worker.py --out /tmp/n2_$1_$2 --exchange $1 --graph $2 > /tmp/log_$1_$2 2>&1 || tail -5 /tmp/log_$1_$2
done
python - << 'PY'
import numpy as np
ref = np.load("/tmp/n1.npz")
for name in ("peer_0", "peer_1", "nccl_0"):
    g = np.load(f"/tmp/n2_{name}.npz")
    print(name, str(g["exchange"]), str(g["transport"]), "status", int(g["status"]))
    for t in range(len(ref["stats"])):
        a, b = ref["stats"][t], g["stats"][t]
        print("  step", t, "LL", a[0], b[0], "rel", abs(a[0]-b[0])/abs(a[0]), "| loss rel", abs(a[3]-b[3])/abs(a[3]))
    print("  params rel l2", np.linalg.norm(ref["params"]-g["params"])/np.linalg.norm(ref["params"]))
PY
