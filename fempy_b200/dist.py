"""Element-partitioned assembly across the GPUs of one box.

The reference has no distributed code, so the only compatibility target is the final global CSR:

* nodes are owned in contiguous global-id ranges; rank r owns DOF rows [2*n0_r, 2*n1_r) and returns that row
  block with GLOBAL column indices, so stacking the rank blocks reproduces the single-GPU pattern bit for bit;
* each rank runs the single-GPU kernel on a LOCAL mesh made of every element that touches one of its owned nodes,
  so the pattern AND the values of the owned rows are complete on the rank that owns them.

Two ways to treat the elements that straddle a partition boundary:

``mode="halo"`` (default, no data-path collective): every rank evaluates all elements of its local mesh, i.e. a
  boundary element is evaluated by each rank owning one of its nodes (one element layer per boundary: 0.2% extra
  work for a 512-wide strip).  The owner-computes rule of the row-owner kernel then finishes every owned row
  locally; the rows of halo nodes are incomplete and simply not returned.  Nothing is exchanged at assembly time.

``mode="exchange"``: every element is assigned to exactly one rank -- the owner of its lowest-numbered node --
  and gets thickness 0 elsewhere.  Contributions to rows of a higher rank ("interface" partial sums of K and R)
  are sent to the owner: one batched send/recv per neighbour over NCCL, then an indexed add.  The slot positions
  of the interface entries inside the owner's row block are negotiated once, at plan time, by exchanging
  (row, column) DOF pairs.  Kept for meshes whose boundary layers are thick (high-order or 3D-like adjacency).

The local assembler is injectable (``assembler=``): the GPU path uses ``GpuLocalAssembler`` (AssemblyPlan on the
rank's device); the CPU tests drive the same partition / exchange logic over gloo with an oracle-based assembler.
"""
import numpy as np


def node_ranges(numNodes, world):
    """Contiguous node ownership: bounds[r] .. bounds[r+1] (as even as possible)."""
    base, extra = divmod(numNodes, world)
    sizes = np.full(world, base, dtype=np.int64)
    sizes[:extra] += 1
    return np.concatenate(([0], np.cumsum(sizes)))


class Partition:
    """Host-side description of one rank's share of the mesh (pure numpy)."""

    def __init__(self, numNodes, connectivity, rank, world, mode="halo"):
        if mode not in ("halo", "exchange"):
            raise ValueError("mode must be 'halo' or 'exchange'")
        self.mode = mode
        self.rank, self.world, self.numNodes = rank, world, int(numNodes)
        self.bounds = node_ranges(numNodes, world)
        self.n0, self.n1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.elemIds = {}   # type -> global element ids (within the type) of the local mesh
        self.active = {}    # type -> bool mask over elemIds: True = evaluated by this rank
        nodeSets = [np.arange(self.n0, self.n1, dtype=np.int64)]
        for elType, conn in connectivity.items():
            conn = np.asarray(conn, dtype=np.int64)
            touches = ((conn >= self.n0) & (conn < self.n1)).any(axis=1)
            ids = np.flatnonzero(touches)
            owner = np.searchsorted(self.bounds, conn[ids].min(axis=1), side="right") - 1
            self.elemIds[elType] = ids
            self.active[elType] = np.ones(ids.size, dtype=bool) if mode == "halo" else owner == rank
            nodeSets.append(conn[ids].reshape(-1))
        # local numbering = rank order of the global ids: monotone, so sorted CSR columns stay sorted
        self.localNodes = np.unique(np.concatenate(nodeSets))
        self.ownedLo = int(np.searchsorted(self.localNodes, self.n0))
        self.numOwned = self.n1 - self.n0
        self.localConn = {t: np.searchsorted(self.localNodes, np.asarray(connectivity[t], dtype=np.int64)[ids])
                          for t, ids in self.elemIds.items()}
        self.numLocalNodes = int(self.localNodes.size)
        self.numActiveElems = int(sum(m.sum() for m in self.active.values()))

    def local_thickness(self, thickness, typeOffsets):
        """Thickness of the local mesh's elements; in exchange mode 0 for the elements another rank evaluates."""
        out = []
        for t, ids in self.elemIds.items():
            th = np.asarray(thickness, dtype=np.float64)[typeOffsets[t] + ids].copy()
            th[~self.active[t]] = 0.0
            out.append(th)
        return np.concatenate(out) if out else np.zeros(0)

    def owned_dof_slice(self):
        return slice(2 * self.ownedLo, 2 * (self.ownedLo + self.numOwned))

    def global_dofs(self, localDofs):
        localDofs = np.asarray(localDofs, dtype=np.int64)
        return 2 * self.localNodes[localDofs // 2] + localDofs % 2

    def local_dofs(self, globalDofs):
        globalDofs = np.asarray(globalDofs, dtype=np.int64)
        return 2 * np.searchsorted(self.localNodes, globalDofs // 2) + globalDofs % 2


class DistributedAssembly:
    """Per-rank driver: local fused assembly (+ interface exchange in exchange mode) -> this rank's CSR row block.

    assembler must provide:
        pattern() -> (indptr, indices) of the LOCAL mesh (int64, local DOF numbering)
        assemble(thickness, states, loads, bcDOF, bcVal, applyBCs) -> (Kdata, R) as torch tensors on `device`
    group: torch.distributed process group (or None for world size 1)
    """

    def __init__(self, partition, assembler, device, group=None):
        import torch

        self.torch = torch
        self.part, self.asm, self.device, self.group = partition, assembler, device, group
        indptr, indices = assembler.pattern()
        self.indptr, self.indices = np.asarray(indptr), np.asarray(indices)
        p = partition
        rows = p.owned_dof_slice()
        self.row0, self.row1 = rows.start, rows.stop
        self.k0, self.k1 = int(self.indptr[self.row0]), int(self.indptr[self.row1])
        # row block with global column ids (what the caller stacks into the global matrix)
        self.blockIndptr = (self.indptr[self.row0 : self.row1 + 1] - self.k0).astype(np.int64)
        self.blockIndices = p.global_dofs(self.indices[self.k0 : self.k1])
        self._negotiate()
        self._bcFilter = None
        self._ownedMask = None

    # -- plan time: who sends what where -------------------------------------------------------------------
    def _negotiate(self):
        import torch.distributed as dist

        p = self.part
        # interface rows: DOF rows of nodes above my range that my ACTIVE elements touch (none in halo mode:
        # the owner evaluates those elements itself)
        ghostNodes = set()
        for t, conn in ({} if p.mode == "halo" else p.localConn).items():
            act = conn[p.active[t]]
            ghostNodes.update(np.unique(act[act >= p.ownedLo + p.numOwned]).tolist())
        ghostNodes = np.array(sorted(ghostNodes), dtype=np.int64)
        ghostOwner = np.searchsorted(p.bounds, p.localNodes[ghostNodes], side="right") - 1 if ghostNodes.size else np.zeros(0, dtype=np.int64)
        self.sendK, self.sendR, outgoing = {}, {}, {}
        for s in np.unique(ghostOwner):
            nodes = ghostNodes[ghostOwner == s]
            dofRows = np.stack((2 * nodes, 2 * nodes + 1), axis=1).reshape(-1)
            starts, ends = self.indptr[dofRows], self.indptr[dofRows + 1]
            pos = np.concatenate([np.arange(a, b) for a, b in zip(starts, ends)]) if dofRows.size else np.zeros(0, dtype=np.int64)
            rowOfPos = np.repeat(dofRows, ends - starts)
            self.sendK[int(s)] = pos.astype(np.int64)
            self.sendR[int(s)] = dofRows.astype(np.int64)
            outgoing[int(s)] = (p.global_dofs(rowOfPos), p.global_dofs(self.indices[pos]), p.global_dofs(dofRows))
        # every rank learns what it will receive (tiny: interface entries only)
        if p.world > 1 and p.mode != "halo":
            gathered = [None] * p.world
            dist.all_gather_object(gathered, outgoing, group=self.group)
        else:
            gathered = [outgoing]
        self.recvK, self.recvR = {}, {}
        keys = None
        for src, out in enumerate(gathered):
            if src == p.rank or p.rank not in out:
                continue
            rowsG, colsG, rdofG = out[p.rank]
            if keys is None:  # (row, col) -> position in my local CSR; CSR order is sorted by (row, col)
                rowOf = np.repeat(np.arange(self.indptr.size - 1, dtype=np.int64), np.diff(self.indptr))
                keys = rowOf * (2 * p.numLocalNodes) + self.indices
            want = p.local_dofs(rowsG) * (2 * p.numLocalNodes) + p.local_dofs(colsG)
            pos = np.searchsorted(keys, want)
            if not np.array_equal(keys[pos], want):
                raise RuntimeError("interface entry outside the owner's pattern (inconsistent partition)")
            self.recvK[src] = pos.astype(np.int64)
            self.recvR[src] = p.local_dofs(rdofG).astype(np.int64)
        t = self.torch
        dev = self.device
        self._sendKd = {s: t.from_numpy(v).to(dev) for s, v in self.sendK.items()}
        self._sendRd = {s: t.from_numpy(v).to(dev) for s, v in self.sendR.items()}
        self._recvKd = {s: t.from_numpy(v).to(dev) for s, v in self.recvK.items()}
        self._recvRd = {s: t.from_numpy(v).to(dev) for s, v in self.recvR.items()}
        self._recvKeep = {}

    def interface_bytes(self):
        """bytes sent per assembly by this rank (K + R partial sums)"""
        return 8 * int(sum(v.size for v in self.sendK.values()) + sum(v.size for v in self.sendR.values()))

    def set_bcs(self, bcDOFGlobal, bcValues):
        """BCs are applied by the owner of the row only; received partial sums never touch a constrained row."""
        p = self.part
        bcDOFGlobal = np.asarray(bcDOFGlobal, dtype=np.int64)
        bcValues = np.asarray(bcValues, dtype=np.float64)
        mine = (bcDOFGlobal // 2 >= p.n0) & (bcDOFGlobal // 2 < p.n1)
        self.bcLocal = p.local_dofs(bcDOFGlobal[mine])
        self.bcVal = bcValues[mine]
        constrained = np.zeros(2 * p.numLocalNodes, dtype=bool)
        constrained[self.bcLocal] = True
        rowOf = np.repeat(np.arange(self.indptr.size - 1, dtype=np.int64), np.diff(self.indptr))
        t = self.torch
        self._recvKeep = {}
        for src in self.recvK:
            keepK = ~constrained[rowOf[self.recvK[src]]]
            keepR = ~constrained[self.recvR[src]]
            self._recvKeep[src] = (t.from_numpy(keepK).to(self.device), t.from_numpy(keepR).to(self.device))

    # -- assembly time -----------------------------------------------------------------------------------------
    def assemble(self, thicknessLocal, statesLocal, loadsLocal=None, applyBCs=True):
        """Local fused assembly, interface exchange, owned row block.  Returns (Kblock, Rblock) torch tensors
        (views into the local arrays): CSR values of the owned DOF rows and the owned residual entries."""
        import torch.distributed as dist

        t = self.torch
        bcd = getattr(self, "bcLocal", None) if applyBCs else None
        bcv = getattr(self, "bcVal", None) if applyBCs else None
        if loadsLocal is not None and self.part.mode == "exchange":
            # loads are applied by the OWNER of a row only: the partial sums of ghost rows travel to their owner, which
            # subtracts its own loads -- entries given for ghost rows are dropped here (halo mode never returns those rows)
            if t.is_tensor(loadsLocal):
                if self._ownedMask is None:
                    m = np.zeros(2 * self.part.numLocalNodes)
                    m[self.part.owned_dof_slice()] = 1.0
                    self._ownedMask = t.from_numpy(m).to(loadsLocal.device)
                loadsLocal = loadsLocal * self._ownedMask
            else:
                masked = np.zeros(2 * self.part.numLocalNodes)
                masked[self.part.owned_dof_slice()] = np.asarray(loadsLocal, dtype=np.float64)[self.part.owned_dof_slice()]
                loadsLocal = masked
        K, R = self.asm.assemble(thicknessLocal, statesLocal, loadsLocal, bcd, bcv, applyBCs)
        ops, recvBufs = [], {}
        for s in self._sendKd:
            buf = t.cat((K.index_select(0, self._sendKd[s]), R.index_select(0, self._sendRd[s])))
            ops.append(dist.P2POp(dist.isend, buf, s, group=self.group))
        for s in self._recvKd:
            recvBufs[s] = t.empty(self._recvKd[s].numel() + self._recvRd[s].numel(), dtype=K.dtype, device=K.device)
            ops.append(dist.P2POp(dist.irecv, recvBufs[s], s, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        for s, buf in recvBufs.items():
            nk = self._recvKd[s].numel()
            bk, br = buf[:nk], buf[nk:]
            if applyBCs and s in self._recvKeep:
                keepK, keepR = self._recvKeep[s]
                bk, br = bk * keepK, br * keepR
            K.index_add_(0, self._recvKd[s], bk)
            R.index_add_(0, self._recvRd[s], br)
        return K[self.k0 : self.k1], R[self.row0 : self.row1]


class GpuLocalAssembler:
    """The single-GPU fused kernel on the rank's local mesh (AssemblyPlan on the rank's device)."""

    def __init__(self, partition, elementTables, coords, material, device):
        import torch

        from .plan import AssemblyPlan

        self.torch, self.part, self.material = torch, partition, material
        self.device = torch.device("cuda", device)
        self.localCoords = np.ascontiguousarray(np.asarray(coords, dtype=np.float64)[partition.localNodes])
        self.plan = AssemblyPlan(partition.numLocalNodes, 2, 2, device=device)
        for elType, conn in partition.localConn.items():
            N, dN, w = elementTables[elType]
            self.plan.addBlock(N, dN, w, conn)
        self.plan.finalize()
        self.d_coords = torch.from_numpy(self.localCoords).to(self.device)
        self.d_K = torch.empty(self.plan.nnz, dtype=torch.float64, device=self.device)
        self.d_R = torch.empty(self.plan.numDOF, dtype=torch.float64, device=self.device)

    def pattern(self):
        return self.plan.pattern()

    def to_device(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    def assemble(self, thickness, states, loads, bcDOF, bcVal, applyBCs):
        t = self.torch
        th = thickness if t.is_tensor(thickness) else self.to_device(thickness)
        st = states if t.is_tensor(states) else self.to_device(states)
        ld = None if loads is None else (loads if t.is_tensor(loads) else self.to_device(loads))
        if applyBCs and bcDOF is not None:
            self.plan.setBCs(bcDOF, bcVal)  # femb_plan_set_bcs returns early when the lists are the ones it already holds
        self.plan.setStream(t.cuda.current_stream(self.device).cuda_stream)
        self.plan.assembleDev(self.d_coords, st, th, self.material, bool(applyBCs and bcDOF is not None and len(bcDOF)), ld, self.d_K, self.d_R)
        return self.d_K, self.d_R
