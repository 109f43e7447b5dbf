"""CPU emulation of the multi-GPU data flow from the host plans (tacs_b200/csrc/plan.cpp).

Each rank computes element matrices for its own elements with the oracle, lays them out in the staging
order of its plan, ships the off-rank rows according to the plan's exchange lists through a transport
(in-process dictionary, or torch.distributed/gloo between real processes), runs the gather plans and
the SpMV with the column halo, and the result is compared with the serial oracle assembly.
"""
import numpy as np

from tests import oracle_port

KIND_NN = {1: 4, 2: 9, 3: 8, 4: 27}


class LocalTransport:
    """All ranks live in this process: send() stores, recv() looks up."""

    def __init__(self):
        self.box = {}

    def send(self, src, dst, tag, arr):
        self.box[(src, dst, tag)] = np.array(arr, copy=True)

    def recv(self, src, dst, tag, shape):
        return self.box[(src, dst, tag)].reshape(shape)


def upper_index(nn, i, j):
    return i * nn - i * (i - 1) // 2 + (j - i)


def rank_staging(plan, mesh, kind, desc, new_nodes, u_global):
    """Element matrices / residuals of this rank's elements in staging layout (+ empty tail for received rows):
    upper node pairs (i <= j) only; pairs with a direct target go straight into the value array [Aloc | Bext]."""
    s = plan.scalars()
    nn = KIND_NN[kind]
    nu = nn * (nn + 1) // 2
    bs = 6 if kind <= 2 else 3
    Ke = np.zeros((s["local_blocks"] + s["recv_blocks"], bs, bs))
    Re = np.zeros((s["local_node_slots"] + s["recv_node_slots"], bs))
    nnz = plan.array("Aloc_rowp")[-1] + (plan.array("Bext_rowp")[-1] if plan.array("Bext_rowp").size else 0)
    Avals = np.full((nnz, bs, bs), np.nan)  # every block must be written exactly once (directly or by the gather)
    written = np.zeros(nnz, np.int32)
    dmap = plan.array("dmap").reshape(-1, nn * nn)
    X = np.zeros((mesh["num_nodes"], 3))
    X[new_nodes] = mesh["Xpts"]
    conn_g = plan.array("elem_conn_global").reshape(-1, nn)
    assert s["local_blocks"] == s["nelems"] * nu
    for e in range(s["nelems"]):  # single element family: staging order == local element order
        nodes = conn_g[e]
        uu = u_global.reshape(-1, bs)[nodes].ravel()
        res, mat = oracle_port.element(kind, desc, X[nodes].ravel(), uu)
        blk = mat.reshape(nn, bs, nn, bs).transpose(0, 2, 1, 3)
        for i in range(nn):
            for j in range(nn):
                d = dmap[e, i * nn + j]
                if d >= 0:
                    assert dmap[e, j * nn + i] >= 0, "direct targets come in mirror pairs"
                    Avals[d] = blk[i, j]
                    written[d] += 1
                elif i <= j:
                    Ke[e * nu + upper_index(nn, i, j)] = blk[i, j]
        Re[e * nn:(e + 1) * nn] = res.reshape(nn, bs)
    return Ke, Re, Avals, written


def exchange(plans, rank, transport, name, src_of, dst_of, tag):
    """Post this rank's sends of exchange `name`; returns a closure that completes the receives."""
    P = plans[rank]
    peers, ptr, idx = P.array(name + "_send_peers"), P.array(name + "_send_ptr"), P.array(name + "_send_idx")
    for k, p in enumerate(peers):
        transport.send(rank, int(p), tag, src_of[idx[ptr[k]:ptr[k + 1]]])

    def finish():
        rpeers, rptr = P.array(name + "_recv_peers"), P.array(name + "_recv_ptr")
        for k, p in enumerate(rpeers):
            n = rptr[k + 1] - rptr[k]
            dst_of(int(rptr[k]), n, transport.recv(int(p), rank, tag, (n,) + src_of.shape[1:]))

    return finish


def gather_sum(ptr, src, Re):
    out = np.zeros((ptr.size - 1,) + Re.shape[1:])
    for b in range(ptr.size - 1):
        for k in range(ptr[b], ptr[b + 1]):  # ascending element order, like the device kernel
            out[b] += Re[src[k]]
    return out


def gather_blocks(P, Ke, Avals, written):
    """Replay of gather_blocks*_kernel: block gb_blk[g] sums its sources 2*slot + transposed flag in list order."""
    blk, ptr, src = P.array("gb_blk"), P.array("gb_ptr"), P.array("gb_src")
    for g in range(blk.size):
        acc = np.zeros(Ke.shape[1:])
        for k in range(ptr[g], ptr[g + 1]):
            v = Ke[src[k] >> 1]
            acc += v.T if (src[k] & 1) else v
        Avals[blk[g]] = acc
        written[blk[g]] += 1
    assert np.all(written == 1), "every block is written exactly once"
    assert int((written == 1).sum()) - blk.size == P.scalars()["direct_blocks"]


def apply_bcs(plan, bc_global, A_vals, B_vals, res, u_owned, bs):
    lo, hi = plan.array("owner_range")[[plan.rank_, plan.rank_ + 1]]
    np_ = plan.scalars()["np"]
    ar, ac = plan.array("Aloc_rowp"), plan.array("Aloc_cols")
    br = plan.array("Bext_rowp")
    for g in bc_global:
        if lo <= g < hi:
            r = g - lo
            for k in range(ar[r], ar[r + 1]):
                A_vals[k] = 0.0
                if ac[k] == r:
                    A_vals[k] = np.eye(bs)
            if r >= np_:
                B_vals[br[r - np_]:br[r - np_ + 1]] = 0.0
            res[r] = u_owned[r]


def run_rank_phase1(plans, rank, transport, mesh, kind, desc, new_nodes, u_global):
    P = plans[rank]
    bs = 6 if kind <= 2 else 3
    Ke, Re, Avals, written = rank_staging(P, mesh, kind, desc, new_nodes, u_global)
    s = P.scalars()

    def put_blocks(off, n, data):
        Ke[s["local_blocks"] + off:s["local_blocks"] + off + n] = data

    def put_rows(off, n, data):
        Re[s["local_node_slots"] + off:s["local_node_slots"] + off + n] = data

    f1 = exchange(plans, rank, transport, "blocks", Ke, put_blocks, "K")
    f2 = exchange(plans, rank, transport, "rows", Re, put_rows, "R")
    return dict(Ke=Ke, Re=Re, Avals=Avals, written=written, finish=(f1, f2), bs=bs)


def run_rank_phase2(plans, rank, transport, st, bc_global, u_global, x_global):
    P = plans[rank]
    bs = st["bs"]
    for f in st["finish"]:
        f()
    gather_blocks(P, st["Ke"], st["Avals"], st["written"])
    nnzA = P.array("Aloc_rowp")[-1]
    A, Bv = st["Avals"][:nnzA], st["Avals"][nnzA:]
    res = gather_sum(P.array("r_ptr"), P.array("r_src"), st["Re"])
    lo, hi = P.array("owner_range")[[rank, rank + 1]]
    P.rank_ = rank
    apply_bcs(P, bc_global, A, Bv, res, u_global.reshape(-1, bs)[lo:hi], bs)
    # column halo for the SpMV
    x_owned = x_global.reshape(-1, bs)[lo:hi]
    x_ext = np.zeros((P.array("ext_col_nodes").size, bs))

    def put_cols(off, n, data):
        x_ext[off:off + n] = data

    st["spmv_finish"] = exchange(plans, rank, transport, "cols", x_owned, put_cols, "X")
    st.update(A=A, B=Bv, res=res, x_owned=x_owned, x_ext=x_ext)
    return st


def run_rank_phase3(plans, rank, st):
    P = plans[rank]
    st["spmv_finish"]()
    ar, ac = P.array("Aloc_rowp"), P.array("Aloc_cols")
    br, bc = P.array("Bext_rowp"), P.array("Bext_cols")
    np_ = P.scalars()["np"]
    y = np.zeros_like(st["x_owned"])
    for r in range(ar.size - 1):
        for k in range(ar[r], ar[r + 1]):
            y[r] += st["A"][k] @ st["x_owned"][ac[k]]
    for r in range(br.size - 1):
        for k in range(br[r], br[r + 1]):
            y[np_ + r] += st["B"][k] @ st["x_ext"][bc[k]]
    st["y"] = y
    return st


def check_against_serial(plans, rank, st, serial):
    """Every owned block / residual entry / SpMV entry equals the serial assembly in the same numbering."""
    P = plans[rank]
    lo, hi = P.array("owner_range")[[rank, rank + 1]]
    ar, ac = P.array("Aloc_rowp"), P.array("Aloc_cols")
    br, bc = P.array("Bext_rowp"), P.array("Bext_cols")
    ext_cols = P.array("ext_col_nodes")
    np_ = P.scalars()["np"]
    srow, scol, sA = serial["rowp"], serial["cols"], serial["A"]
    scale = np.abs(sA).max()
    worst = 0.0
    for r in range(hi - lo):
        cols = [lo + c for c in ac[ar[r]:ar[r + 1]]]
        vals = [st["A"][k] for k in range(ar[r], ar[r + 1])]
        if r >= np_:
            cols += [ext_cols[c] for c in bc[br[r - np_]:br[r - np_ + 1]]]
            vals += [st["B"][k] for k in range(br[r - np_], br[r - np_ + 1])]
        order = np.argsort(cols)
        g = lo + r
        assert np.array_equal(np.asarray(cols)[order], scol[srow[g]:srow[g + 1]]), "row pattern differs from serial"
        for m, o in enumerate(order):
            worst = max(worst, np.abs(vals[o] - sA[srow[g] + m]).max() / scale)
    bs = st["bs"]
    e_res = np.abs(st["res"].ravel() - serial["res"][bs * lo:bs * hi]).max() / max(np.abs(serial["res"]).max(), 1e-300)
    e_y = np.abs(st["y"].ravel() - serial["y"][bs * lo:bs * hi]).max() / np.abs(serial["y"]).max()
    return worst, e_res, e_y
