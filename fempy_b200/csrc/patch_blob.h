// Layout of a patch "program blob": everything one CTA needs to assemble the CSR rows of a patch of nodes,
// packed contiguously (16-byte aligned sections) so that a single cp.async.bulk brings it into shared memory.
//
//   header   8 x int32   ne, nOwned, nLocal, nnb, numColours[0..3] (colours of the scatter phase of local node a)
//   LNODE    nLocal x int32     global node id of each local node; the first nOwned are the owned nodes in patch order
//   NODE     nOwned x 8 bytes   {int32 off = nrowptr[i]; uint16 start = first node-block (patch-relative);
//                                uint8 deg; uint8 diagK = position of i in its own row}
//   NBROW    nnb x uint16       owned index r of each (owned node, neighbour) block, in accumulator order
//   EINFO    ne x 32 bytes      per element, for each local node a: uint16 rowBase[a] (accumulator offset of the
//                               node's row pair, in double2 units), uint8 deg[a], uint8 colour[a] (0xFF: the node
//                               is not owned by this patch), then uint8 k[a][b] = position of node b in row a
//   CONN     nn*ne x uint16     local node ids, node-major: [a*ne + t]
//   ELEM     ne x int32         global element id (thickness lookup)
//
// The shared-memory accumulators of a patch mirror the global CSR layout: owned node r (nbStart = NODE.start,
// deg) owns 2*deg double2 at double2 offset 2*start + r (one pad double2 per preceding node makes the row stride
// an odd multiple of 16 bytes: conflict-free 128-bit accesses from neighbouring elements): DOF row 0 =
// [k0.xx k0.xy | k1.xx k1.xy | ...], then DOF row 1.  A whole
// node's accumulators are therefore one contiguous 32*deg-byte run that maps to K + 4*off: the write-out is a
// straight 16-byte-per-thread copy, global double2 index = 2*off + j for shared index 2*start + r + j.
#pragma once

struct BlobLayout {
    int offLNODE, offNODE, offNBROW, offEINFO, offCONN, offELEM, bytes;
};

__host__ __device__ inline int pad16(int b) { return (b + 15) & ~15; }

__host__ __device__ inline BlobLayout blob_layout(int nn, int ne, int nOwned, int nLocal, int nnb) {
    BlobLayout L;
    L.offLNODE = 32;
    L.offNODE = L.offLNODE + pad16(4 * nLocal);
    L.offNBROW = L.offNODE + pad16(8 * nOwned);
    L.offEINFO = L.offNBROW + pad16(2 * nnb);
    L.offCONN = L.offEINFO + 32 * ne;
    L.offELEM = L.offCONN + pad16(2 * nn * ne);
    L.bytes = L.offELEM + pad16(4 * ne);
    return L;
}

struct __align__(16) ElemInfo {  // 32 bytes, element families with at most 4 nodes
    unsigned short rowBase[4];
    unsigned char deg[4];
    unsigned char colour[4];
    unsigned char k[16];
};
