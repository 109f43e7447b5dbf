// Host-side (CPU only, no CUDA) planning of one rank's share of the mesh: the integer pipeline that
// fixes the local element order, node maps, matrix sparsity and the atomic-free gather plans, and
// the neighbour exchange lists that replace the reference's MPI traffic.
//
// What it reproduces (bit-exact integer results are the parity target):
//   TACSCreator::createTACS local selection          src/TACSCreator.cpp:493-747 (stable sort by partition)
//   TACSAssembler::computeExtNodes / local order     src/TACSAssembler.cpp:1013-1098, 1681-1770
//   TACSAssembler::createMat + TACSMatDistribute     src/TACSAssembler.cpp:3384-3440,
//       (Aloc / Bext patterns incl. rows contributed   src/bpmat/TACSMatDistribute.cpp:65-415, 453-692
//        by elements of other ranks, np, ext columns)
//   TACSBVecDistribute plans (state halo, column halo)  src/bpmat/TACSBVecDistribute.cpp:299-467
//   TACSMatDistribute off-rank rows                   src/bpmat/TACSMatDistribute.cpp:1154-1267
//
// Every rank holds the whole mesh (synthetic meshes are generated on every rank; the Python layer
// broadcasts a root-only mesh), so each plan is computed locally and deterministically: both ends
// of an exchange derive the same ordered lists and no set-up communication is needed.
#pragma once

#include <memory>
#include <vector>

namespace tb2 {

// the whole mesh in the creator's NEW node numbering (after first-touch renumbering)
struct GlobalMesh {
  int num_nodes = 0, num_elements = 0, size = 1;
  std::vector<int> ptr, conn;    // element -> nodes
  std::vector<int> part;         // element -> owning rank
  std::vector<int> owner_range;  // size+1, contiguous node ranges
  int ownerOf(int node) const;
};

// One neighbour exchange at a fixed granularity (a "chunk" of doubles): the sender packs
// source[send_idx[k]] chunks for each peer, the receiver gets each peer's chunks contiguously.
struct ExchangePlan {
  std::vector<int> send_peers, send_ptr, send_idx;  // send_ptr has send_peers.size()+1 entries
  std::vector<int> recv_peers, recv_ptr;            // offsets (chunks) inside the receive region
  int sendTotal() const { return send_ptr.empty() ? 0 : send_ptr.back(); }
  int recvTotal() const { return recv_ptr.empty() ? 0 : recv_ptr.back(); }
};

struct HostBCSR {
  int bsize = 0, nrows = 0, ncols = 0;
  std::vector<int> rowp, cols;
  long nnzb() const { return rowp.empty() ? 0 : rowp[nrows]; }
};

// Matrix staging keeps the node pairs (i, j) with i <= j of an element only (the element matrices of the path are
// symmetric: K = sum B^T C B with a symmetric C, likewise the mass matrix): nn(nn+1)/2 slots per element, row-major
// over the upper triangle. The block of the directed pair (k, j) with k > j is the transpose of slot (j, k).
inline int upper_pairs(int nn) { return nn * (nn + 1) / 2; }
inline int upper_index(int nn, int i, int j) { return i * nn - i * (i - 1) / 2 + (j - i); }  // i <= j
// Node pairs the element kernel of a family looks up a direct target for (HostPlan::dmap): the low-order kernels
// test only the pairs that are alone in their block on a regular mesh -- the diagonal of a quadrilateral, the body
// diagonals of a hexahedron (node i against nn-1-i in the tensor-product node order) -- so that their store code
// stays cheap; the quadratic families look up every pair. kind = ElemKind (kernels.h).
inline bool direct_candidate(int kind, int nn, int i, int j) {
  return (kind == 1 || kind == 3) ? (i + j == nn - 1) : true;
}

// Blocks shipped to the owner of some of an element's nodes (bit set `mask`): the upper pairs that touch a masked
// node, in ascending upper_index order. Both ends of the exchange evaluate these two functions.
int shipped_pairs(int nn, unsigned mask);
int shipped_rank(int nn, unsigned mask, int i, int j);  // position of pair (i <= j) inside the shipped set

// one (element, local node) pair that contributes to an owned matrix/residual row
struct RowContribution {
  int gelem;        // global element id (summation order key)
  int nn;           // nodes of the element
  const int *conn;  // its connectivity in global (new) numbering
  int res_slot;     // residual staging slot (units of one node block)
  int slot_base;    // first matrix staging slot of the element (local) / of its shipped set (received)
  int lelem;        // local element index, or -1 when the element belongs to another rank
  unsigned mask;    // received contributions: the element's nodes owned by this rank (defines the shipped set)
  // staging source of the directed pair (k, j) of this element: 2 * slot + (1 when the slot holds the transpose)
  int source(int k, int j) const {
    const int i0 = k < j ? k : j, j0 = k < j ? j : k;
    const int off = lelem >= 0 ? upper_index(nn, i0, j0) : shipped_rank(nn, mask, i0, j0);
    return 2 * (slot_base + off) + (k > j ? 1 : 0);
  }
};

struct HostPlan {
  int bs = 0, rank = 0, size = 1;
  std::vector<int> owner_range;
  // local mesh (TACSAssembler's element order and numbering)
  int nelems = 0, nowned = 0, nlocal = 0, ext_before = 0, ext_after = 0;
  std::vector<int> elem_global, elem_ptr, elem_conn_global, elem_conn_local;
  std::vector<int> ext_nodes;
  // staging layout: elements grouped by kernel family keep their local order inside a group
  std::vector<int> elem_kind;                // per local element (ElemKind)
  std::vector<int> group_kinds;              // distinct kinds in first-appearance order
  std::vector<std::vector<int>> group_elems; // local element ids of each group
  std::vector<long> group_block_base, group_node_base, group_pair_base;
  std::vector<long> elem_block_base, elem_node_base, elem_pair_base;  // per local element (pair_base: units of nn^2)
  long local_pairs = 0;                               // sum of nn^2 over the local elements
  long local_blocks = 0, local_node_slots = 0;        // staging produced by this rank's element kernels
  long recv_blocks = 0, recv_node_slots = 0;          // staging received from other ranks (tail region)
  // contributions to owned rows, ascending (owned row, global element)
  std::vector<int> adj_ptr;
  std::vector<RowContribution> adj;
  std::vector<int> adj_i;  // local node index i of each contribution
  // residual gather plan
  std::vector<int> r_ptr, r_src;
  // matrix (filled by buildMatrix)
  bool has_matrix = false;
  HostBCSR Aloc, Bext;
  int np = 0;
  std::vector<int> ext_col_nodes;
  // Matrix values live in one array [Aloc blocks | Bext blocks]; "block index" below is an index into it.
  // direct map: for the directed node pair (k, j) of local element e, dmap[elem_pair_base[e] + k*nn + j] is the
  // block index the element kernel writes straight into -- possible when that block and its mirror (j, k) each have
  // exactly one contribution and both rows are owned here -- or -1 (the pair goes through the staging area).
  std::vector<int> dmap;
  // gather plan of all other blocks: block gb_blk[g] sums the staging sources gb_src[gb_ptr[g] .. gb_ptr[g+1])
  // (RowContribution::source encoding) in ascending global element order -- the order of the reference's serial loop
  std::vector<int> gb_blk, gb_ptr, gb_src;
  static constexpr int kGatherBucketShift = 6;  // gathered blocks are ordered by (last staging slot >> 6)
  std::vector<int> gb_bucket_start;             // first gathered block of every bucket of 64 staging slots
  long gatherEnd(long slot) const;
  long direct_blocks = 0;   // blocks written by the element kernels
  long staged_blocks = 0;   // upper node-pair blocks the local element kernels write to the staging area
  // neighbour exchanges (empty on one rank)
  ExchangePlan state;   // chunk = one node block of a state vector: owned node -> ext slots of peers
  ExchangePlan cols;    // chunk = one node block of x: owned node -> x_ext of peers (SpMV)
  ExchangePlan blocks;  // chunk = one bs x bs staging block: rows owned by a peer
  ExchangePlan rows;    // chunk = one node block of the residual staging
  std::shared_ptr<const GlobalMesh> gm;

  int localNode(int global) const;
  int globalNode(int local) const;
  // `elem_kinds[e]` is the kernel family of global element e (same on every rank)
  int build(std::shared_ptr<const GlobalMesh> mesh, int bs, int rank, const std::vector<int> &elem_kinds);
  int buildMatrix();
};

int kind_nodes(int kind);

}  // namespace tb2
