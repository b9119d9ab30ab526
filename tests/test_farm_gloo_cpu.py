"""CPU suite: the N > 1 plumbing (shard, search locally, gather records) under gloo with world_size 2.
The local searcher is the oracle here -- what is under test is the sharding/gather logic, which is
the same code the NCCL run uses."""
import os
import socket

import numpy as np
import pytest

from flydog_sdr_gps_b200 import sats as S, scenarios, synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleEngine:
    """Stand-in for AcqEngine in the farm classes: same search_device(ptr, ptr, n, sel, stream) call, the oracle as the
    searcher, CPU tensors as the 'device'.  What is under test is the sharding / padding / gather / reassembly."""

    def __init__(self, table):
        from oracle import oracle_py as O
        self.O, self.table, self.device = O, table, "cpu"

    def search_device(self, packed_ptr, out_ptr, n_cap, sel=None, stream_ptr=None):
        import ctypes as C
        n_sel = len(self.table) if sel is None else len(sel)
        packed = np.ctypeslib.as_array((C.c_uint8 * (n_cap * 8192)).from_address(packed_ptr)).reshape(n_cap, 8192)
        out = np.ctypeslib.as_array((C.c_uint8 * (n_cap * n_sel * 24)).from_address(out_ptr))
        rec = np.stack([self.O.search(c, self.table, sel=sel, nthreads=1) for c in packed])
        out[:] = rec.reshape(-1).view(np.uint8)


def _worker(rank, world, port, mode, caps, out_dir):
    import torch.distributed as dist
    from flydog_sdr_gps_b200 import farm
    from oracle import oracle_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table = S.navstar()[:6]
    if mode == "capture":
        fn = lambda c: np.stack([O.search(x, table, nthreads=1) for x in c])
        rec = farm.search_sharded_by_capture(fn, caps, len(table), dist=dist)
    elif mode == "sat":
        fn = lambda c, sel: O.search(c, table, sel=sel, nthreads=1)
        rec = farm.search_sharded_by_sat(fn, caps[0], len(table), dist=dist)
    elif mode == "capture_farm":   # 3 captures over 2 ranks: uneven shards, padded gather
        f = farm.CaptureFarm(_OracleEngine(table), len(caps), 8192, len(table), dist=dist, device="cpu")
        f.load(caps)
        rec = f.search().copy()
        assert np.array_equal(f.search_resident().numpy(), f.h_all.numpy())
    else:                          # 5 satellites over 2 ranks
        f = farm.SatFarm(_OracleEngine(table[:5]), 5, 8192, dist=dist, device="cpu")
        f.load(caps[0])
        rec = f.search().copy()
    np.save(os.path.join(out_dir, "%s_%d.npy" % (mode, rank)), rec)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["capture", "sat", "capture_farm", "sat_farm"])
def test_world_size_2_gather_equals_single_process(tmp_path, mode, oracle):
    import torch.multiprocessing as mp
    table = S.navstar()[:6]
    caps = np.stack([synth.make_capture(s, 1, table, [(s % 6, 400 * s, 250.0 * s, 48, 0.1)]) for s in range(1, 4)])
    port = _free_port()
    mp.spawn(_worker, args=(2, port, mode, caps, str(tmp_path)), nprocs=2, join=True)
    if mode.startswith("capture"):
        want = np.stack([oracle.search(c, table, nthreads=1) for c in caps])
    elif mode == "sat":
        want = oracle.search(caps[0], table, nthreads=1)
    else:
        want = oracle.search(caps[0], table[:5], nthreads=1)
    for r in range(2):
        got = np.load(os.path.join(str(tmp_path), "%s_%d.npy" % (mode, r)))
        assert got.shape == want.shape
        assert got.tobytes() == want.tobytes()
