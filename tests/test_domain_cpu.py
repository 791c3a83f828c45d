"""CPU: single-mesh domain decomposition (finite-volume-graph-network_b200/domain.py) — the partitioner, the
exchange plans and the per-GnBlock orchestration (partial node sums of interface nodes, ghost-cell x', ghost next-uv), run with ALL parts
in one process (LocalComm) and the ORACLE's torch arithmetic as the backend, against the single-domain oracle rollout.  The GPU test
(tests/test_gpu_domain.py) runs the same orchestration on the kernels, one part per GPU over NCCL."""
import numpy as np
import pytest
import torch
import torch.nn.functional as Fnn

import fvgn_b200 as fv
from fvgn_b200 import domain as D
from oracle import fvgn_oracle as O
from tests import helpers as H


class OracleBackend:
    def __init__(self, orc):
        self.o = orc

    def build_agg(self, agg_dst, n_rows):
        return agg_dst.long(), n_rows

    def encode(self, pr, ms):
        o, C = self.o, pr.p.C_own
        uv = pr.x[:C, 2:4]
        cell_in = o.norm["_cell_normalizer."](uv, False)
        edge_in = o.norm["_edge_normalizer."](o._edge_feats(pr.ea, pr.ftl[:, 0:1], 9), False)
        return {"x": O.mlp(o.sd, "model.encoder.cb_encoder.", cell_in), "e": O.mlp(o.sd, "model.encoder.eb_encoder.", edge_in), "uv": uv}

    def node_partial(self, pr, s, layer):
        dst, n = pr.agg
        twoway = torch.cat(torch.chunk(s["e"], 2, dim=-1), dim=0)
        return torch.zeros((n, 64), dtype=s["e"].dtype).index_add_(0, dst, twoway)

    def cell_block(self, pr, s, na, layer, ms):
        cn = pr.cnT.long()
        agg = (na[cn[0]] + na[cn[1]] + na[cn[2]]) / 3.0
        xp = O.mlp(self.o.sd, f"model.processer_list.{layer}.cb_module.net.", torch.cat((s["x"], agg), dim=-1))
        buf = torch.zeros((pr.p.C_own + pr.p.G, 128), dtype=xp.dtype)
        buf[:pr.p.C_own] = xp
        s["x"] = s["x"] + xp
        return buf

    def edge_block(self, pr, s, xp, layer, ms):
        nc = pr.nc.long()
        en = O.mlp(self.o.sd, f"model.processer_list.{layer}.eb_module.net.", torch.cat((xp[nc[0]], xp[nc[1]], s["e"]), dim=1))
        s["e"] = s["e"] + en

    def decode_integrate(self, pr, s, ms):
        o, nc = self.o, pr.nc.long()
        pred = O.mlp(o.sd, "model.decoder.edge_decode_module.", s["e"], layer_norm=False)
        raw = (pr.ftl[:, 1:2] * (ms["dt"] / ((pr.area[nc[0]] + pr.area[nc[1]]) / 2))).view(-1, 1)
        fa = Fnn.batch_norm(raw, o.bn["running_mean"], o.bn["running_var"], o.bn["weight"], o.bn["bias"], False, 0.1, 1e-5)
        _, du = O.integrator(pred[:, 0:2], pred[:, 2:3], pred[:, 3:5], pr.unv, ms["rho"], 1.0, fa, pr.cfT.long(), 0)
        return o.norm["_face_uvp_flux_output_normalizer."].inverse(pred[:, 0:3]), o.norm["_acc_output_normalizer."].inverse(du) + s["uv"]

    def next_graph(self, pr, uv_all):
        nc = pr.nc.long()
        pr.x = torch.cat((pr.x[:, 0:2], uv_all), dim=1)
        euv = uv_all[nc[0]] - uv_all[nc[1]]
        interior = (pr.ftl[:, 0] == O.NORMAL) | (pr.ftl[:, 0] == O.OUTFLOW)
        euv[~interior] = pr.bc[~interior]
        pr.ea = torch.cat((euv, pr.ea[:, 2:5]), dim=1)


def _single_domain(m, tab, steps):
    on, oe, oc = O.build_graphs(tab, m["velocity"], m["pressure"])
    on, oe, oc, mi, mb = O.valid_datapreprocessing(on, oe, oc, 0)
    x0, ea0, bc = oc.x.clone(), oc.edge_attr.clone(), oe.y[:, 0, 0:2].clone()
    orc = O.FVGNOracle(H.torch_reference_state_dict(0), H.params()).eval()
    ct, Re, rmp = oc.x[:, 0:1].clone(), oc.x[:, 1:2].clone(), oc.edge_attr[:, 2:5].clone()
    out = []
    with torch.no_grad():
        for _ in range(steps):
            uvp, nuv = orc(oc, oe, on, H.RHO, H.MU, H.DT, 9, 0)
            out.append((uvp.clone(), nuv.clone()))
            O.create_next_graph(oc, oe, mb, nuv, ct, Re, rmp)
    return x0, ea0, bc, out


@pytest.mark.parametrize("n_parts", [2, 3, 5])
def test_partition_tables_and_exchange_plans(n_parts):
    m = fv.synthetic.make("cyl", seed=2, T=3, n_target=900)
    tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    tab["mesh_pos"], tab["node_type"] = m["mesh_pos"], m["node_type"]
    parts, part_of = D.decompose(tab, n_parts)
    C, F = tab["cells_node"].shape[0], tab["face"].shape[1]
    assert sorted(np.concatenate([p.own for p in parts]).tolist()) == list(range(C))          # every cell owned exactly once
    assert max(p.C_own for p in parts) - min(p.C_own for p in parts) <= 1                      # balanced
    owners = np.zeros(F, dtype=np.int64)
    for p in parts:
        assert (p.neighbour_cell >= 0).all() and (p.cells_face_T >= 0).all() and (p.cells_node_T >= 0).all()
        # local tables map back to the global ones
        cells = np.concatenate([p.own, p.ghost])
        assert np.array_equal(cells[p.neighbour_cell], tab["neighbour_cell"][:, p.faces])
        assert np.array_equal(p.faces[p.cells_face_T.T], tab["cells_face"][p.own])
        assert np.array_equal(p.nodes[p.cells_node_T.T], tab["cells_node"][p.own])
        owned_faces = p.agg_dst.reshape(2, -1)[0] < p.N
        owners[p.faces[owned_faces]] += 1
        for q, ix in p.halo_send.items():   # what I send to q is what q expects from me, in the same (global) order
            assert np.array_equal(p.own[ix], parts[q].ghost[parts[q].halo_recv[p.rank] - parts[q].C_own])
        for q, ix in p.node_share.items():
            assert np.array_equal(p.nodes[ix], parts[q].nodes[parts[q].node_share[p.rank]])
    assert (owners == 1).all()                                                                  # every face aggregated by exactly one part
    # receive-buffer layouts of the peer-memory exchanges (csrc/halo.cu): every rank derives every rank's layout from `parts`; a sender's
    # count equals the receiver's region, regions are 256-byte aligned, disjoint, inside the two parity copies, behind the flag block
    lays = [D.exchange_layout(parts, r) for r in range(n_parts)]
    for p, lay in zip(parts, lays):
        spans = [(0, len(D.EXCHANGE_KINDS) * n_parts * 8)]
        for kind, rb in D.EXCHANGE_KINDS:
            k = lay[kind]
            sends = p.node_share if kind == "node" else p.halo_send
            for q in sends:
                assert lays[q][kind]["n"][p.rank] == len(sends[q]), (kind, p.rank, q)
            for q, n in k["n"].items():
                assert k["off"][q] % 256 == 0 and k["off"][q] + n * rb <= k["stride"]
                for parity in (0, 1):
                    spans.append((k["base"] + parity * k["stride"] + k["off"][q], k["base"] + parity * k["stride"] + k["off"][q] + n * rb))
        spans.sort()
        assert all(a[1] <= b[0] for a, b in zip(spans[:-1], spans[1:])) and spans[-1][1] <= lay["total"]


@pytest.mark.parametrize("kind,n_parts", [("cyl", 2), ("cyl", 3), ("grid", 4)])
def test_decomposed_rollout_equals_the_single_domain_rollout(kind, n_parts):
    m = fv.synthetic.make(kind, seed=4, T=3, **({"n_target": 700} if kind == "cyl" else {"nx": 21, "ny": 13}))
    tab = O.build_connectivity(m["mesh_pos"], m["cells"], m["node_type"], "B")
    tab["mesh_pos"], tab["node_type"] = m["mesh_pos"], m["node_type"]
    steps = 3
    x0, ea0, bc, want = _single_domain(m, tab, steps)
    parts, _ = D.decompose(tab, n_parts)
    orc = O.FVGNOracle(H.torch_reference_state_dict(0), H.params()).eval()
    runs = [D.PartRollout(p, OracleBackend(orc), *D.initial_part_state(p, x0, ea0, bc), device="cpu") for p in parts]
    with torch.no_grad():
        outs = D.rollout_steps(runs, D.LocalComm(n_parts), {"n_layers": 15, "dt": H.DT, "rho": H.RHO}, steps)
    C, F = x0.shape[0], ea0.shape[0]
    for k in range(steps):
        nuv = torch.zeros((C, 2))
        uvp = torch.zeros((F, 3))
        seen = torch.zeros(F, dtype=torch.long)
        for p, o in zip(parts, outs):
            nuv[torch.as_tensor(p.own)] = o[k][1]
            f = torch.as_tensor(p.faces)
            # a cut face lives on two parts: both computed the same values
            dup = seen[f] > 0
            assert torch.equal(uvp[f][dup], o[k][0][dup])
            uvp[f] = o[k][0]
            seen[f] += 1
        e1, e2 = H.rel_l2(nuv, want[k][1]), H.rel_l2(uvp, want[k][0])
        assert e1 < 2e-6 * (k + 1) and e2 < 2e-6 * (k + 1), (k, e1, e2)
