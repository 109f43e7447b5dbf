// Host-only planning (see plan.h). No CUDA in this file: the CPU test-suite exercises it directly
// through the tacsb200_plan_* entry points, including 2-rank runs over gloo.
#include "plan.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <map>
#include <thread>

namespace tb2 {

int kind_nodes(int kind) {
  switch (kind) {
    case 1: return 4;
    case 2: return 9;
    case 3: return 8;
    case 4: return 27;
  }
  return 0;
}

int shipped_pairs(int nn, unsigned mask) {
  const int rest = nn - __builtin_popcount(mask);
  return upper_pairs(nn) - upper_pairs(rest);
}

int shipped_rank(int nn, unsigned mask, int i, int j) {
  int r = 0;
  for (int a = 0; a < i; a++) {
    if (mask & (1u << a)) r += nn - a;                           // every pair (a, a..nn-1) touches the mask
    else r += __builtin_popcount(mask >> a);                     // pairs (a, b > a) with b masked
  }
  if (mask & (1u << i)) r += j - i;
  else r += __builtin_popcount((mask >> i) & ((1u << (j - i)) - 1u));
  return r;
}

int GlobalMesh::ownerOf(int node) const {
  return (int)(std::upper_bound(owner_range.begin(), owner_range.end(), node) - owner_range.begin()) - 1;
}

static void plan_parallel_for(long n, const std::function<void(long, long)> &fn) {
  unsigned hw = std::thread::hardware_concurrency();
  long nt = hw ? hw : 4;
  if (nt > 64) nt = 64;
  if (n < 4096 || nt <= 1) {
    fn(0, n);
    return;
  }
  std::vector<std::thread> th;
  long chunk = (n + nt - 1) / nt;
  for (long t = 0; t < nt; t++) {
    long lo = t * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back(fn, lo, hi);
  }
  for (auto &t : th) t.join();
}

// TacsUniqueSort (src/TacsUtilities.cpp:76-99): ascending, negatives dropped, duplicates removed
static int unique_sort(std::vector<int> &a) {
  std::sort(a.begin(), a.end());
  size_t i = 0;
  while (i < a.size() && a[i] < 0) i++;
  size_t n = 0;
  for (; i < a.size(); i++)
    if (n == 0 || a[n - 1] != a[i]) a[n++] = a[i];
  a.resize(n);
  return (int)n;
}

// TACSAssembler::getLocalNodeNum (src/TACSAssembler.cpp:1681-1731)
int HostPlan::localNode(int g) const {
  const int lo = owner_range[rank], hi = owner_range[rank + 1];
  if (g >= lo && g < hi) return ext_before + (g - lo);
  auto it = std::lower_bound(ext_nodes.begin(), ext_nodes.end(), g);
  if (it == ext_nodes.end() || *it != g) return -1;
  int k = (int)(it - ext_nodes.begin());
  return k < ext_before ? k : nowned + k;
}

// TACSAssembler::getGlobalNodeNum (src/TACSAssembler.cpp:1733-1770)
int HostPlan::globalNode(int l) const {
  if (l < ext_before) return ext_nodes[l];
  if (l < ext_before + nowned) return owner_range[rank] + (l - ext_before);
  return ext_nodes[l - nowned];
}

struct RemoteRow {
  int gelem, i;
};

int HostPlan::build(std::shared_ptr<const GlobalMesh> mesh, int _bs, int _rank, const std::vector<int> &elem_kinds) {
  gm = mesh;
  bs = _bs;
  rank = _rank;
  size = mesh->size;
  owner_range = mesh->owner_range;
  const int lo = owner_range[rank], hi = owner_range[rank + 1];
  nowned = hi - lo;
  const GlobalMesh &g = *mesh;

  // local elements: ascending global id inside the partition (compare_arg_sort, TACSCreator.cpp:37-47)
  elem_ptr.assign(1, 0);
  for (int e = 0; e < g.num_elements; e++) {
    if (g.part[e] != rank) continue;
    elem_global.push_back(e);
    elem_kind.push_back(elem_kinds[e]);
    for (int k = g.ptr[e]; k < g.ptr[e + 1]; k++) elem_conn_global.push_back(g.conn[k]);
    elem_ptr.push_back((int)elem_conn_global.size());
  }
  nelems = (int)elem_global.size();

  // external nodes (TACSAssembler::computeExtNodes, :1013-1098)
  ext_nodes.clear();
  for (int n : elem_conn_global)
    if (n < lo || n >= hi) ext_nodes.push_back(n);
  unique_sort(ext_nodes);
  ext_before = (int)(std::lower_bound(ext_nodes.begin(), ext_nodes.end(), lo) - ext_nodes.begin());
  ext_after = (int)ext_nodes.size() - ext_before;
  nlocal = nowned + (int)ext_nodes.size();
  elem_conn_local.resize(elem_conn_global.size());
  for (size_t k = 0; k < elem_conn_global.size(); k++) elem_conn_local[k] = localNode(elem_conn_global[k]);

  // staging layout by kernel family
  {
    std::map<int, int> gidx;
    for (int e = 0; e < nelems; e++) {
      int kind = elem_kind[e];
      if (!gidx.count(kind)) {
        gidx[kind] = (int)group_kinds.size();
        group_kinds.push_back(kind);
        group_elems.emplace_back();
      }
      group_elems[gidx[kind]].push_back(e);
    }
    elem_block_base.assign(nelems, 0);
    elem_node_base.assign(nelems, 0);
    elem_pair_base.assign(nelems, 0);
    local_blocks = local_node_slots = local_pairs = 0;
    for (size_t gi = 0; gi < group_kinds.size(); gi++) {
      const long nn = kind_nodes(group_kinds[gi]), nu = upper_pairs((int)nn);
      group_block_base.push_back(local_blocks);
      group_node_base.push_back(local_node_slots);
      group_pair_base.push_back(local_pairs);
      for (size_t k = 0; k < group_elems[gi].size(); k++) {
        elem_block_base[group_elems[gi][k]] = local_blocks + (long)k * nu;
        elem_node_base[group_elems[gi][k]] = local_node_slots + (long)k * nn;
        elem_pair_base[group_elems[gi][k]] = local_pairs + (long)k * nn * nn;
      }
      local_blocks += (long)group_elems[gi].size() * nu;
      local_node_slots += (long)group_elems[gi].size() * nn;
      local_pairs += (long)group_elems[gi].size() * nn * nn;
    }
  }

  // interface scan: what crosses rank boundaries
  std::vector<std::vector<int>> state_need(size), cols_need(size), rows_send(size), blocks_send(size);
  std::vector<std::vector<RemoteRow>> remote(size);
  if (size > 1) {
    std::vector<int> owners;
    int local_index = 0;
    for (int e = 0; e < g.num_elements; e++) {
      const int p = g.part[e], b = g.ptr[e], nn = g.ptr[e + 1] - b;
      owners.resize(nn);
      bool mine = false, other = false;
      for (int k = 0; k < nn; k++) {
        owners[k] = g.ownerOf(g.conn[b + k]);
        if (owners[k] == rank) mine = true;
        else other = true;
      }
      if (p == rank) {
        // rows of nodes owned elsewhere are shipped to their owners (TACSMatDistribute :801-834, 1154-1267):
        // the residual row of each such node, and once per owner the upper node-pair blocks that touch its nodes
        for (int k = 0; k < nn; k++)
          if (owners[k] != rank) {
            rows_send[owners[k]].push_back((int)(elem_node_base[local_index] + k));
            bool first = true;
            for (int m = 0; m < k; m++)
              if (owners[m] == owners[k]) { first = false; break; }
            if (!first) continue;
            unsigned mask = 0;
            for (int m = k; m < nn; m++)
              if (owners[m] == owners[k]) mask |= 1u << m;
            for (int i = 0; i < nn; i++)
              for (int j = i; j < nn; j++)
                if ((mask >> i | mask >> j) & 1u)
                  blocks_send[owners[k]].push_back((int)(elem_block_base[local_index] + upper_index(nn, i, j)));
          }
        local_index++;
      } else if (mine) {
        for (int k = 0; k < nn; k++)
          if (owners[k] == rank) {
            state_need[p].push_back(g.conn[b + k]);  // rank p reads my node as one of its external nodes
            remote[p].push_back({e, k});
          }
      }
      if (mine && other) {
        // every other owner in this element has my nodes as external matrix columns
        for (int k = 0; k < nn; k++) {
          if (owners[k] == rank) continue;
          bool seen = false;
          for (int m = 0; m < k; m++)
            if (owners[m] == owners[k]) { seen = true; break; }
          if (seen) continue;
          for (int m = 0; m < nn; m++)
            if (owners[m] == rank) cols_need[owners[k]].push_back(g.conn[b + m]);
        }
      }
    }
  }
  auto fill_send = [&](ExchangePlan &x, std::vector<std::vector<int>> &lists, bool make_unique, int offset) {
    x.send_peers.clear();
    x.send_ptr.assign(1, 0);
    x.send_idx.clear();
    for (int p = 0; p < size; p++) {
      if (make_unique) unique_sort(lists[p]);
      if (lists[p].empty()) continue;
      x.send_peers.push_back(p);
      for (int v : lists[p]) x.send_idx.push_back(v - offset);
      x.send_ptr.push_back((int)x.send_idx.size());
    }
  };
  fill_send(state, state_need, true, lo);
  fill_send(cols, cols_need, true, lo);
  fill_send(rows, rows_send, false, 0);
  fill_send(blocks, blocks_send, false, 0);
  // receive side of the state halo: my external nodes are sorted, so each owner's are contiguous
  state.recv_peers.clear();
  state.recv_ptr.assign(1, 0);
  for (size_t k = 0; k < ext_nodes.size();) {
    int o = g.ownerOf(ext_nodes[k]);
    size_t m = k;
    while (m < ext_nodes.size() && ext_nodes[m] < owner_range[o + 1]) m++;
    state.recv_peers.push_back(o);
    state.recv_ptr.push_back((int)m);
    k = m;
  }
  // receive side of the off-rank rows: peers ascending, rows in the sender's (element, node) order
  rows.recv_peers.clear();
  rows.recv_ptr.assign(1, 0);
  blocks.recv_peers.clear();
  blocks.recv_ptr.assign(1, 0);
  recv_blocks = recv_node_slots = 0;
  std::vector<RowContribution> remote_rc;
  std::vector<int> remote_node, remote_i;
  for (int p = 0; p < size; p++) {
    if (remote[p].empty()) continue;
    int cur_elem = -1, cur_base = 0;
    unsigned cur_mask = 0;
    for (const RemoteRow &rr : remote[p]) {
      const int b = g.ptr[rr.gelem], nn = g.ptr[rr.gelem + 1] - b;
      if (rr.gelem != cur_elem) {
        // the sender ships the element's blocks once: the upper pairs that touch my nodes (shipped_pairs / _rank)
        cur_elem = rr.gelem;
        cur_mask = 0;
        for (int m = 0; m < nn; m++)
          if (g.conn[b + m] >= lo && g.conn[b + m] < hi) cur_mask |= 1u << m;
        cur_base = (int)(local_blocks + recv_blocks);
        recv_blocks += shipped_pairs(nn, cur_mask);
      }
      RowContribution rc;
      rc.gelem = rr.gelem;
      rc.nn = nn;
      rc.conn = &g.conn[b];
      rc.res_slot = (int)(local_node_slots + recv_node_slots);
      rc.slot_base = cur_base;
      rc.lelem = -1;
      rc.mask = cur_mask;
      remote_rc.push_back(rc);
      remote_node.push_back(g.conn[b + rr.i] - lo);
      remote_i.push_back(rr.i);
      recv_node_slots += 1;
    }
    rows.recv_peers.push_back(p);
    rows.recv_ptr.push_back((int)recv_node_slots);
    blocks.recv_peers.push_back(p);
    blocks.recv_ptr.push_back((int)recv_blocks);
  }
  if (local_blocks + recv_blocks >= (1L << 30) || local_pairs >= (1L << 31)) {
    fprintf(stderr, "[%d] tacs_b200: %ld staging blocks / %ld node pairs exceed the 32-bit gather index\n", rank,
            local_blocks + recv_blocks, local_pairs);
    return 1;
  }

  // contributions to owned rows, ordered by (row, global element): the serial summation order
  {
    std::vector<int> count(nowned + 1, 0);
    for (int e = 0; e < nelems; e++)
      for (int k = elem_ptr[e]; k < elem_ptr[e + 1]; k++) {
        int n = elem_conn_global[k];
        if (n >= lo && n < hi) count[n - lo + 1]++;
      }
    for (int n : remote_node) count[n + 1]++;
    for (int i = 0; i < nowned; i++) count[i + 1] += count[i];
    adj_ptr = count;
    adj.resize(count[nowned]);
    adj_i.resize(count[nowned]);
    std::vector<int> cursor(count.begin(), count.end() - 1);
    for (int e = 0; e < nelems; e++) {
      const int nn = elem_ptr[e + 1] - elem_ptr[e];
      for (int k = 0; k < nn; k++) {
        int n = elem_conn_global[elem_ptr[e] + k];
        if (n < lo || n >= hi) continue;
        RowContribution rc;
        rc.gelem = elem_global[e];
        rc.nn = nn;
        rc.conn = &elem_conn_global[elem_ptr[e]];
        rc.res_slot = (int)(elem_node_base[e] + k);
        rc.slot_base = (int)elem_block_base[e];
        rc.lelem = e;
        rc.mask = 0;
        int pos = cursor[n - lo]++;
        adj[pos] = rc;
        adj_i[pos] = k;
      }
    }
    for (size_t r = 0; r < remote_rc.size(); r++) {
      int pos = cursor[remote_node[r]]++;
      adj[pos] = remote_rc[r];
      adj_i[pos] = remote_i[r];
    }
    if (!remote_rc.empty()) {
      // merge local and remote contributions of each row by global element id
      plan_parallel_for(nowned, [&](long r0, long r1) {
        std::vector<std::pair<int, int>> order;
        std::vector<RowContribution> tmp;
        std::vector<int> tmpi;
        for (long r = r0; r < r1; r++) {
          const int b = adj_ptr[r], n = adj_ptr[r + 1] - b;
          bool sorted = true;
          for (int k = 1; k < n; k++)
            if (adj[b + k].gelem < adj[b + k - 1].gelem) { sorted = false; break; }
          if (sorted) continue;
          order.clear();
          for (int k = 0; k < n; k++) order.emplace_back(adj[b + k].gelem, k);
          std::stable_sort(order.begin(), order.end());
          tmp.assign(adj.begin() + b, adj.begin() + b + n);
          tmpi.assign(adj_i.begin() + b, adj_i.begin() + b + n);
          for (int k = 0; k < n; k++) {
            adj[b + k] = tmp[order[k].second];
            adj_i[b + k] = tmpi[order[k].second];
          }
        }
      });
    }
  }
  r_ptr = adj_ptr;
  r_src.resize(adj.size());
  for (size_t k = 0; k < adj.size(); k++) r_src[k] = adj[k].res_slot;
  return 0;
}

// number of gathered blocks (a prefix of gb_blk) that are complete once every staging slot below `slot` is written
long HostPlan::gatherEnd(long slot) const {
  if (gb_bucket_start.empty()) return 0;
  long b = slot >> kGatherBucketShift;  // buckets [0, b) hold last slots < b * bucket <= slot
  if (b >= (long)gb_bucket_start.size()) b = (long)gb_bucket_start.size() - 1;
  return gb_bucket_start[b];
}

int HostPlan::buildMatrix() {
  if (has_matrix) return 0;
  const int lo = owner_range[rank], hi = owner_range[rank + 1];
  // sorted unique columns of every owned row (TACSAssembler::computeLocalNodeToNodeCSR :1899-2053 +
  // TacsSortAndUniquifyCSR; rows contributed by other ranks are merged as TACSMatDistribute does)
  std::vector<int> rowlen(nowned, 0);
  auto row_columns = [&](long r, std::vector<int> &buf) {
    buf.clear();
    for (int p = adj_ptr[r]; p < adj_ptr[r + 1]; p++)
      for (int j = 0; j < adj[p].nn; j++) buf.push_back(adj[p].conn[j]);
    unique_sort(buf);
  };
  plan_parallel_for(nowned, [&](long r0, long r1) {
    std::vector<int> buf;
    for (long r = r0; r < r1; r++) {
      row_columns(r, buf);
      rowlen[r] = (int)buf.size();
    }
  });
  std::vector<long> gptr(nowned + 1, 0);
  for (int r = 0; r < nowned; r++) gptr[r + 1] = gptr[r] + rowlen[r];
  if (gptr[nowned] >= (1L << 31)) {
    fprintf(stderr, "[%d] tacs_b200: %ld blocks exceed the reference's 32-bit rowp (BCSRMat.cpp:234)\n", rank,
            gptr[nowned]);
    return 1;
  }
  std::vector<int> gcols(gptr[nowned]);
  plan_parallel_for(nowned, [&](long r0, long r1) {
    std::vector<int> buf;
    for (long r = r0; r < r1; r++) {
      row_columns(r, buf);
      if (!buf.empty()) memcpy(&gcols[gptr[r]], buf.data(), buf.size() * sizeof(int));
    }
  });
  // Aloc: owned columns (local index); Bext: external columns (index into the ascending unique list);
  // np: first owned row that has an external column (TACSMatDistribute.cpp:573-597, 662-680)
  np = nowned;
  for (int r = 0; r < nowned && np == nowned; r++)
    for (long k = gptr[r]; k < gptr[r + 1]; k++)
      if (gcols[k] < lo || gcols[k] >= hi) { np = r; break; }
  ext_col_nodes.clear();
  for (long k = 0; k < gptr[nowned]; k++)
    if (gcols[k] < lo || gcols[k] >= hi) ext_col_nodes.push_back(gcols[k]);
  unique_sort(ext_col_nodes);
  Aloc.bsize = bs; Aloc.nrows = nowned; Aloc.ncols = nowned;
  Aloc.rowp.assign(nowned + 1, 0);
  Aloc.cols.clear();
  Bext.bsize = bs; Bext.nrows = nowned - np; Bext.ncols = (int)ext_col_nodes.size();
  Bext.rowp.assign(Bext.nrows + 1, 0);
  Bext.cols.clear();
  for (int r = 0; r < nowned; r++) {
    for (long k = gptr[r]; k < gptr[r + 1]; k++) {
      int c = gcols[k];
      if (c >= lo && c < hi) Aloc.cols.push_back(c - lo);
      else Bext.cols.push_back((int)(std::lower_bound(ext_col_nodes.begin(), ext_col_nodes.end(), c) -
                                     ext_col_nodes.begin()));
    }
    Aloc.rowp[r + 1] = (int)Aloc.cols.size();
    if (r >= np) Bext.rowp[r - np + 1] = (int)Bext.cols.size();
  }
  // Direct map and gather plan. Block index space: [Aloc blocks | Bext blocks].
  const long nnzA = Aloc.nnzb(), nnzB = Bext.nnzb(), nnz = nnzA + nnzB;
  auto locate = [&](int r, int gcol) -> long {
    if (gcol >= lo && gcol < hi) {
      const int *b = &Aloc.cols[0] + Aloc.rowp[r], *e = &Aloc.cols[0] + Aloc.rowp[r + 1];
      return std::lower_bound(b, e, gcol - lo) - &Aloc.cols[0];
    }
    int c = (int)(std::lower_bound(ext_col_nodes.begin(), ext_col_nodes.end(), gcol) - ext_col_nodes.begin());
    const int *b = &Bext.cols[0] + Bext.rowp[r - np], *e = &Bext.cols[0] + Bext.rowp[r - np + 1];
    return nnzA + (std::lower_bound(b, e, c) - &Bext.cols[0]);
  };
  // pass 1: contributions per block; target block of every directed pair of the local elements. A block belongs to
  // one row, so the row-parallel loop has a single writer per counter and per dmap entry.
  std::vector<int> cnt(nnz, 0);
  dmap.assign(local_pairs, -1);
  plan_parallel_for(nowned, [&](long r0, long r1) {
    for (long r = r0; r < r1; r++)
      for (int p = adj_ptr[r]; p < adj_ptr[r + 1]; p++) {
        const RowContribution &rc = adj[p];
        const int k = adj_i[p];
        for (int j = 0; j < rc.nn; j++) {
          const long t = locate((int)r, rc.conn[j]);
          cnt[t]++;
          if (rc.lelem >= 0) dmap[elem_pair_base[rc.lelem] + (long)k * rc.nn + j] = (int)t;
        }
      }
  });
  // pass 2: a pair is written directly when its block and the mirror block are fed by this element alone.
  // TACSB200_DIRECT_KINDS (bit k-1 = element kind k, default all) limits this to some families (measurements).
  unsigned direct_kinds = 0xffffffffu;
  if (const char *env = getenv("TACSB200_DIRECT_KINDS")) direct_kinds = (unsigned)strtoul(env, nullptr, 0);
  std::vector<unsigned char> is_direct(nnz, 0);
  plan_parallel_for(nelems, [&](long e0, long e1) {
    for (long e = e0; e < e1; e++) {
      const int nn = elem_ptr[e + 1] - elem_ptr[e];
      const bool kind_on = (direct_kinds >> (elem_kind[e] - 1)) & 1u;
      int *dm = &dmap[elem_pair_base[e]];
      for (int k = 0; k < nn; k++)
        for (int j = k; j < nn; j++) {
          const int t = dm[k * nn + j], t2 = dm[j * nn + k];
          const bool direct = kind_on && t >= 0 && t2 >= 0 && cnt[t] == 1 && cnt[t2] == 1 &&
                              direct_candidate(elem_kind[e], nn, k, j);
          if (direct) {
            is_direct[t] = 1;
            is_direct[t2] = 1;
          } else {
            dm[k * nn + j] = -1;
            dm[j * nn + k] = -1;
          }
        }
    }
  });
  // pass 3: gather lists of the remaining blocks, sources in (row, ascending global element, k, j) order.
  // The gathered blocks are processed in the order of their LAST staging slot (bucketed by kGatherBucket slots, block
  // order inside a bucket), not in matrix order: (1) a block and its mirror read the same slots and so sit next to
  // each other -- every staging slot comes from DRAM once; (2) the walk streams the staging area front to back; and
  // (3) every block whose last slot lies below a slot s is complete once the elements that own the slots below s
  // have been evaluated: the gather of a chunk of elements can run while the next chunk is being computed
  // (gatherEnd). Blocks fed by received slots (the tail region) come last, after the exchange.
  std::vector<int> first_slot(nnz, -1);  // (holds the last slot)
  plan_parallel_for(nowned, [&](long r0, long r1) {
    for (long r = r0; r < r1; r++)
      for (int p = adj_ptr[r]; p < adj_ptr[r + 1]; p++) {
        const RowContribution &rc = adj[p];
        const int k = adj_i[p];
        for (int j = 0; j < rc.nn; j++) {
          const long t = locate((int)r, rc.conn[j]);
          const int slot = rc.source(k, j) >> 1;
          if (slot > first_slot[t]) first_slot[t] = slot;
        }
      }
  });
  std::vector<int> gidx(nnz, -1);
  gb_blk.clear();
  gb_ptr.assign(1, 0);
  direct_blocks = 0;
  {
    const int kBucketShift = kGatherBucketShift;
    const long nbuckets = ((local_blocks + recv_blocks) >> kBucketShift) + 2;
    std::vector<int> bstart(nbuckets + 1, 0);
    long ngather = 0;
    for (long t = 0; t < nnz; t++) {
      if (is_direct[t]) { direct_blocks++; continue; }
      bstart[(first_slot[t] >> kBucketShift) + 1]++;
      ngather++;
    }
    for (long b = 0; b < nbuckets; b++) bstart[b + 1] += bstart[b];
    gb_bucket_start = bstart;  // first gathered block of every bucket (before the fill below advances bstart)
    gb_blk.resize(ngather);
    for (long t = 0; t < nnz; t++) {
      if (is_direct[t]) continue;
      gb_blk[bstart[first_slot[t] >> kBucketShift]++] = (int)t;
    }
    gb_ptr.resize(ngather + 1);
    for (long g = 0; g < ngather; g++) {
      gidx[gb_blk[g]] = (int)g;
      gb_ptr[g + 1] = gb_ptr[g] + cnt[gb_blk[g]];
    }
  }
  std::vector<int>().swap(first_slot);
  staged_blocks = 0;
  for (int e = 0; e < nelems; e++) {
    const int nn = elem_ptr[e + 1] - elem_ptr[e];
    const int *dm = &dmap[elem_pair_base[e]];
    for (int k = 0; k < nn; k++)
      for (int j = k; j < nn; j++)
        if (dm[k * nn + j] < 0) staged_blocks++;
  }
  gb_src.assign((size_t)gb_ptr.back() + 4, 0);  // + 4: the gather kernels prefetch four sources per list
  {
    std::vector<int> cur(gb_ptr.begin(), gb_ptr.end() - 1);
    plan_parallel_for(nowned, [&](long r0, long r1) {
      for (long r = r0; r < r1; r++)
        for (int p = adj_ptr[r]; p < adj_ptr[r + 1]; p++) {
          const RowContribution &rc = adj[p];
          const int k = adj_i[p];
          for (int j = 0; j < rc.nn; j++) {
            const int gi = gidx[locate((int)r, rc.conn[j])];
            if (gi >= 0) gb_src[cur[gi]++] = rc.source(k, j);
          }
        }
    });
  }
  // receive side of the column halo: ext_col_nodes is sorted, each owner's columns are contiguous
  cols.recv_peers.clear();
  cols.recv_ptr.assign(1, 0);
  for (size_t k = 0; k < ext_col_nodes.size();) {
    int o = gm->ownerOf(ext_col_nodes[k]);
    size_t m = k;
    while (m < ext_col_nodes.size() && ext_col_nodes[m] < owner_range[o + 1]) m++;
    cols.recv_peers.push_back(o);
    cols.recv_ptr.push_back((int)m);
    k = m;
  }
  has_matrix = true;
  return 0;
}

}  // namespace tb2
