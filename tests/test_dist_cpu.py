"""Host-side logic of the element-partitioned multi-GPU path (fempy_b200/dist.py) on CPU: world_size 2 and 3
over gloo, with an oracle-based local assembler standing in for the CUDA kernel.  Checks that the stacked rank
row blocks reproduce the single-process pattern bit for bit and the values to 1e-12."""
import os
import socket

import numpy as np
import pytest

from util import MAT, oracle_blocks


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class OracleLocalAssembler:
    """fem_oracle on the rank's local mesh (thickness 0 marks the elements another rank evaluates)."""

    def __init__(self, partition, coords):
        from oracle import fem_oracle as fo

        self.fo, self.part = fo, partition
        self.blocks = oracle_blocks(partition.localConn)
        self.coords = np.asarray(coords)[partition.localNodes]
        self.ip, self.ix = fo.structural_pattern(self.blocks, partition.numLocalNodes, 2)

    def pattern(self):
        return self.ip, self.ix

    def assemble(self, thickness, states, loads, bcDOF, bcVal, applyBCs):
        import torch

        fo = self.fo
        bc = bcDOF if (applyBCs and bcDOF is not None and len(bcDOF)) else None
        K = fo.assemble_matrix(self.blocks, self.coords, states, thickness, 0, MAT["E"], MAT["nu"], bc, dropZeros=False)
        K.sort_indices()
        assert np.array_equal(K.indptr, self.ip) and np.array_equal(K.indices, self.ix)
        R = fo.assemble_residual(self.blocks, self.coords, states, thickness, 0, MAT["E"], MAT["nu"], loads, bc,
                                 bcVal if bc is not None else None)
        return torch.from_numpy(K.data.copy()), torch.from_numpy(R.copy())


def _mesh(kind):
    from fempy_b200 import meshes

    if kind == "quad":
        return meshes.quad4_grid(9, 7)
    if kind == "triangle":
        return meshes.tri3_grid(7, 6)
    coords, cq = meshes.quad4_grid(8, 5)  # mixed: two element types in one model
    conn = cq["quad"]
    col = np.tile(np.arange(8), 5)
    rest = conn[col >= 4]
    return coords, {"quad": conn[col < 4], "triangle": np.vstack((rest[:, [0, 1, 2]], rest[:, [0, 2, 3]]))}


def _worker(rank, world, port, kind, applyBCs, mode):
    import torch
    import torch.distributed as dist

    from fempy_b200.dist import DistributedAssembly, Partition
    from oracle import fem_oracle as fo

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        coords, conn = _mesh(kind)
        N = coords.shape[0]
        rng = np.random.default_rng(5)
        states = 1e-3 * rng.random((N, 2))
        offsets, e0 = {}, 0
        for t, c in conn.items():
            offsets[t] = e0
            e0 += c.shape[0]
        thick = rng.random(e0) * 1e-2 + 1e-3
        loads = rng.random(2 * N)
        left = np.flatnonzero(coords[:, 0] == 0.0)
        bcDOF = np.concatenate((2 * left, 2 * left + 1, [2 * (N - 1), 2 * (N - 1)]))  # incl. a duplicate on the last node
        bcVal = rng.random(bcDOF.size) * 1e-4

        part = Partition(N, conn, rank, world, mode=mode)
        assert part.numOwned == part.n1 - part.n0 and np.array_equal(part.localNodes[part.ownedLo : part.ownedLo + part.numOwned], np.arange(part.n0, part.n1))
        asm = OracleLocalAssembler(part, coords)
        da = DistributedAssembly(part, asm, torch.device("cpu"))
        if applyBCs:
            da.set_bcs(bcDOF, bcVal)
        # the loads of ALL local nodes, ghost rows included: a row's load is applied by its owner only (the exchange
        # mode drops the ghost entries itself, the halo mode never returns those rows)
        loadsLocal = loads[part.global_dofs(np.arange(2 * part.numLocalNodes))]
        Kb, Rb = da.assemble(part.local_thickness(thick, offsets), states[part.localNodes], loadsLocal, applyBCs)

        gathered = [None] * world
        dist.all_gather_object(gathered, (da.blockIndptr, da.blockIndices, Kb.numpy().copy(), Rb.numpy().copy(), part.numActiveElems, da.interface_bytes()))
        if rank == 0:
            blocks = oracle_blocks(conn)
            Kref = fo.assemble_matrix(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], bcDOF if applyBCs else None, dropZeros=False)
            Kref.sort_indices()
            Rref = fo.assemble_residual(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], loads, bcDOF if applyBCs else None,
                                        bcVal if applyBCs else None)
            indptr = np.concatenate([[0]] + [g[0][1:] + sum(len(h[1]) for h in gathered[:i]) for i, g in enumerate(gathered)])
            indices = np.concatenate([g[1] for g in gathered])
            data = np.concatenate([g[2] for g in gathered])
            R = np.concatenate([g[3] for g in gathered])
            if mode == "exchange":
                assert sum(g[4] for g in gathered) == e0  # every element evaluated exactly once
                assert all(g[5] > 0 for g in gathered[:-1])  # every rank but the last sends interface rows upward
            else:
                assert sum(g[4] for g in gathered) > e0  # boundary elements are evaluated by every rank that needs them
                assert all(g[5] == 0 for g in gathered)  # and nothing is exchanged
            np.testing.assert_array_equal(indptr, Kref.indptr)  # bit-exact global pattern
            np.testing.assert_array_equal(indices, Kref.indices)
            assert np.abs(data - Kref.data).max() < 1e-12 * np.abs(Kref.data).max()
            assert np.abs(R - Rref).max() < 1e-12 * np.abs(Rref).max()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["quad", "triangle", "mixed"])
@pytest.mark.parametrize("applyBCs", [True, False])
@pytest.mark.parametrize("mode", ["halo", "exchange"])
def test_partitioned_assembly_matches_single_process(world, kind, applyBCs, mode):
    if mode == "exchange" and world == 3 and not applyBCs:
        pytest.skip("covered by the other exchange cases (keeps the CPU suite short)")
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(world, _free_port(), kind, applyBCs, mode), nprocs=world, join=True)


def test_partition_properties():
    from fempy_b200 import meshes
    from fempy_b200.dist import Partition, node_ranges

    coords, conn = meshes.quad4_grid(10, 10)
    N = coords.shape[0]
    np.testing.assert_array_equal(node_ranges(10, 3), [0, 4, 7, 10])
    owners = np.zeros(100, dtype=int)
    for world in (1, 2, 4, 8):
        total = 0
        for r in range(world):
            p = Partition(N, conn, r, world, mode="exchange")
            total += p.numActiveElems
            h = Partition(N, conn, r, world)  # default: halo mode evaluates every element touching an owned node
            assert h.mode == "halo" and h.active["quad"].all() and h.numActiveElems == h.elemIds["quad"].size
            ids = p.elemIds["quad"][p.active["quad"]]
            owners[ids] += 1
            assert np.all(np.diff(p.localNodes) > 0)  # monotone local numbering
            np.testing.assert_array_equal(p.global_dofs(p.local_dofs(np.arange(2 * p.n0, 2 * p.n1))), np.arange(2 * p.n0, 2 * p.n1))
        assert total == 100
    assert np.all(owners == 4)  # each element owned exactly once at every world size
