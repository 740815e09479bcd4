"""Multi-process / multi-node launcher (reference scripts/nxdi_distributed_launcher.py:29-155 builds an mpirun(+torchrun)
command with EFA env; on a B200 box it is one ``torchrun`` per node, NCCL over NVLink inside the node and IB/RoCE across)."""
from __future__ import annotations

import argparse
import os
import shlex
import subprocess
import sys
from typing import List


def build_command(nproc_per_node: int, nnodes: int = 1, node_rank: int = 0, master_addr: str = "127.0.0.1", master_port: int = 29500,
                  script: List[str] = ()) -> List[str]:
    return [sys.executable, "-m", "torch.distributed.run", f"--nnodes={nnodes}", f"--node-rank={node_rank}",
            f"--nproc-per-node={nproc_per_node}", "--master-addr", master_addr, "--master-port", str(master_port)] + list(script)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--nproc-per-node", type=int, default=int(os.environ.get("NXDI_GPUS_PER_NODE", "8")))
    ap.add_argument("--nnodes", type=int, default=1)
    ap.add_argument("--node-rank", type=int, default=int(os.environ.get("NODE_RANK", "0")))
    ap.add_argument("--master-addr", default=os.environ.get("MASTER_ADDR", "127.0.0.1"))
    ap.add_argument("--master-port", type=int, default=int(os.environ.get("MASTER_PORT", "29500")))
    ap.add_argument("--dry-run", action="store_true")
    ap.add_argument("command", nargs=argparse.REMAINDER, help="script and its arguments (prefix with --)")
    a = ap.parse_args(argv)
    script = [c for c in a.command if c != "--"]
    cmd = build_command(a.nproc_per_node, a.nnodes, a.node_rank, a.master_addr, a.master_port, script)
    env = dict(os.environ, NCCL_DEBUG=os.environ.get("NCCL_DEBUG", "WARN"))
    print(" ".join(shlex.quote(c) for c in cmd))
    if a.dry_run:
        return 0
    return subprocess.call(cmd, env=env)


if __name__ == "__main__":
    sys.exit(main())
