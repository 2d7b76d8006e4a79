"""System image: the structure-of-arrays form of the reference's config.SystemSpec.

`SystemImage.from_spec` is the host-side equivalent of core.System.SetFromSpec
(reference pkg/core/system.go:82-178): it interns accelerator / type / model / class /
server names to dense indices and resolves the per-server (class, model) lookups that
core.CreateAllocation performs (pkg/core/allocation.go:41-70), so the native side only
ever sees integers and float32 values.
"""
import ctypes as C

import numpy as np

from . import abi


class SystemImage:
    """Host copy of wva_system_soa backed by numpy arrays."""

    def __init__(self, n_servers, n_accels, n_models, n_types):
        self.S, self.A, self.M, self.T = int(n_servers), int(n_accels), int(n_models), int(n_types)
        for name, dt in abi.ACC_FIELDS:
            setattr(self, name, np.zeros(self.A, dtype=dt))
        for name, dt in abi.TYPE_FIELDS:
            setattr(self, name, np.zeros(self.T, dtype=dt))
        for name, dt in abi.PERF_FIELDS:
            setattr(self, name, np.zeros(self.M * self.A, dtype=dt))
        for name, dt in abi.SRV_FIELDS:
            setattr(self, name, np.zeros(self.S, dtype=dt))
        self.srv_cur_acc[:] = abi.ACC_NONE
        self.srv_priority[:] = abi.DEFAULT_PRIORITY
        # optional name tables (only set by from_spec)
        self.acc_names, self.type_names, self.model_names, self.server_names = None, None, None, None

    ALL_FIELDS = abi.ACC_FIELDS + abi.TYPE_FIELDS + abi.PERF_FIELDS + abi.SRV_FIELDS

    def c_struct(self):
        """wva_system_soa pointing into this image's arrays (keep `self` alive while in use)."""
        s = abi.SystemSoa()
        s.n_servers, s.n_accels, s.n_models, s.n_types = self.S, self.A, self.M, self.T
        for name, dt in self.ALL_FIELDS:
            arr = getattr(self, name)
            if arr.dtype != np.dtype(dt) or not arr.flags["C_CONTIGUOUS"]:
                arr = np.ascontiguousarray(arr, dtype=dt)
                setattr(self, name, arr)
            setattr(s, name, abi.ptr(arr, abi._CT[np.dtype(dt)]))
        return s

    def nbytes(self):
        return sum(getattr(self, n).nbytes for n, _ in self.ALL_FIELDS)

    def shard(self, first, count):
        """A new image holding servers [first, first+count) and the full replicated tables."""
        out = SystemImage(count, self.A, self.M, self.T)
        for name, _ in abi.ACC_FIELDS + abi.TYPE_FIELDS + abi.PERF_FIELDS:
            setattr(out, name, getattr(self, name).copy())
        for name, _ in abi.SRV_FIELDS:
            setattr(out, name, getattr(self, name)[first:first + count].copy())
        return out

    def take(self, servers):
        """A new image holding the listed servers (in that order) and the full replicated tables."""
        idx = np.asarray(servers, dtype=np.int64)
        out = SystemImage(len(idx), self.A, self.M, self.T)
        for name, _ in abi.ACC_FIELDS + abi.TYPE_FIELDS + abi.PERF_FIELDS:
            setattr(out, name, getattr(self, name).copy())
        for name, _ in abi.SRV_FIELDS:
            setattr(out, name, getattr(self, name)[idx].copy())
        return out

    # ------------------------------------------------------------------------------
    @classmethod
    def from_spec(cls, spec):
        """Build from a dict shaped like config.SystemSpec's JSON (pkg/config/types.go:11-155).

        Keys follow the reference's json tags: acceleratorData.accelerators[], modelData.models[],
        serviceClassData.serviceClasses[], serverData.servers[], capacityData.count[].
        """
        accs = {}
        for a in spec.get("acceleratorData", {}).get("accelerators", []):      # system.go:91-101
            accs[a["name"]] = a
        acc_names = list(accs)
        acc_idx = {n: i for i, n in enumerate(acc_names)}
        type_names = []
        for a in accs.values():
            t = a.get("type", "")
            if t not in type_names:
                type_names.append(t)
        type_idx = {n: i for i, n in enumerate(type_names)}
        capacity = {}
        for c in spec.get("capacityData", {}).get("count", []):                # system.go:113-122
            capacity[c["type"]] = int(c["count"])

        models = {}
        for pd in spec.get("modelData", {}).get("models", []):                 # system.go:125-134
            models.setdefault(pd["name"], {})[pd["acc"]] = pd                  # model.go:45-54
        model_names = list(models)
        model_idx = {n: i for i, n in enumerate(model_names)}

        classes = {}
        for sc in spec.get("serviceClassData", {}).get("serviceClasses", []):  # system.go:173-177
            prio = int(sc.get("priority", 0))
            if prio < 1 or prio > 100:                                         # serviceclass.go:28-37
                prio = abi.DEFAULT_PRIORITY
            targets = {}
            for mt in sc.get("modelTargets", []) or []:
                targets[mt["model"]] = mt
            classes[sc["name"]] = (prio, targets)

        servers = {}
        for sv in spec.get("serverData", {}).get("servers", []):               # system.go:152-156
            servers[sv["name"]] = sv
        server_names = list(servers)

        img = cls(len(server_names), len(acc_names), len(model_names), len(type_names))
        img.acc_names, img.type_names, img.model_names, img.server_names = acc_names, type_names, model_names, server_names
        for i, n in enumerate(acc_names):
            a = accs[n]
            img.acc_cost[i] = np.float32(a.get("cost", 0.0))
            img.acc_multiplicity[i] = int(a.get("multiplicity", 0))
            img.acc_type[i] = type_idx[a.get("type", "")]
        for i, t in enumerate(type_names):
            img.type_capacity[i] = capacity.get(t, 0)
        for mn, per_acc in models.items():
            m = model_idx[mn]
            for an, pd in per_acc.items():
                if an not in acc_idx:
                    continue   # perf data for an accelerator the system does not have is never consulted
                k = m * img.A + acc_idx[an]
                img.perf_alpha[k] = np.float32(pd.get("decodeParms", {}).get("alpha", 0.0))
                img.perf_beta[k] = np.float32(pd.get("decodeParms", {}).get("beta", 0.0))
                img.perf_gamma[k] = np.float32(pd.get("prefillParms", {}).get("gamma", 0.0))
                img.perf_delta[k] = np.float32(pd.get("prefillParms", {}).get("delta", 0.0))
                img.perf_max_batch[k] = int(pd.get("maxBatchSize", 0))
                img.perf_at_tokens[k] = int(pd.get("atTokens", 0))
                img.perf_acc_count[k] = int(pd.get("accCount", 0))
                img.perf_valid[k] = 1
        for i, n in enumerate(server_names):
            sv = servers[n]
            cls_name = sv.get("class", "") or "Free"                            # server.go:36-39
            model = sv.get("model", "")
            img.srv_model[i] = model_idx.get(model, -1)
            cur = sv.get("currentAlloc", {}) or {}
            load = cur.get("load", {}) or {}
            img.srv_arrival_rpm[i] = np.float32(load.get("arrivalRate", 0.0))
            img.srv_in_tokens[i] = int(load.get("avgInTokens", 0))
            img.srv_out_tokens[i] = int(load.get("avgOutTokens", 0))
            if cls_name in classes:
                prio, targets = classes[cls_name]
                img.srv_priority[i] = prio                                       # server.go:92-97
                if model in targets:
                    mt = targets[model]
                    img.srv_target_valid[i] = 1
                    img.srv_slo_itl[i] = np.float32(mt.get("slo-itl", 0.0))
                    img.srv_slo_ttft[i] = np.float32(mt.get("slo-ttft", 0.0))
                    img.srv_slo_tps[i] = np.float32(mt.get("slo-tps", 0.0))
            img.srv_min_replicas[i] = int(sv.get("minNumReplicas", 0))
            img.srv_max_batch[i] = int(sv.get("maxBatchSize", 0))
            img.srv_keep_acc[i] = 1 if sv.get("keepAccelerator", False) else 0
            cur_acc = cur.get("accelerator", "")
            if cur_acc == "":
                img.srv_cur_acc[i] = abi.ACC_NONE
            else:
                img.srv_cur_acc[i] = acc_idx.get(cur_acc, abi.ACC_UNKNOWN)
            img.srv_cur_replicas[i] = int(cur.get("numReplicas", 0))
            img.srv_cur_cost[i] = np.float32(cur.get("cost", 0.0))
        return img
