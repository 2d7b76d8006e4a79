// wva_host.hpp — C++ host-side mirror of the reference's Go API for the Analyze -> Optimize path,
// implemented over the C-ABI (include/wva_b200.h).  The reference is Go; this image has no Go
// toolchain, so the compiled host layer is C++ and go/ carries the cgo shim written against the
// same C-ABI.  Names, argument meaning and error behaviour follow:
//   pkg/config/types.go            -> wva::config::*Spec
//   pkg/core/{system,server,allocation}.go -> wva::core::System / Server / Allocation
//   pkg/solver/{optimizer,solver}.go       -> wva::solver::Optimizer / Solver
//   pkg/manager/manager.go                 -> wva::manager::Manager
//   internal/modelanalyzer/analyzer.go     -> wva::modelanalyzer::ModelAnalyzer
//   internal/optimizer/optimizer.go        -> wva::optimizer::VariantAutoscalingsEngine
// There is no arithmetic of the path in this file: every number comes back from the CUDA library.
#pragma once

#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/wva_b200.h"

namespace wva {

struct Error : std::runtime_error { int code; Error(int c, const std::string& m) : std::runtime_error(m), code(c) {} };

namespace config {   // pkg/config/types.go:11-155, same field meaning
struct PowerSpec { int Idle = 0, Full = 0, MidPower = 0; float MidUtil = 0; };
struct AcceleratorSpec { std::string Name, Type; int Multiplicity = 0, MemSize = 0, MemBW = 0; PowerSpec Power; float Cost = 0; };
struct AcceleratorCount { std::string Type; int Count = 0; };
struct DecodeParms { float Alpha = 0, Beta = 0; };
struct PrefillParms { float Gamma = 0, Delta = 0; };
struct ModelAcceleratorPerfData { std::string Name, Acc; int AccCount = 0, MaxBatchSize = 0, AtTokens = 0; DecodeParms Decode; PrefillParms Prefill; };
struct ModelTarget { std::string Model; float SLO_ITL = 0, SLO_TTFT = 0, SLO_TPS = 0; };
struct ServiceClassSpec { std::string Name; int Priority = 0; std::vector<ModelTarget> ModelTargets; };
struct ServerLoadSpec { float ArrivalRate = 0; int AvgInTokens = 0, AvgOutTokens = 0; };
struct AllocationData { std::string Accelerator; int64_t NumReplicas = 0, MaxBatch = 0; float Cost = 0, ITLAverage = 0, TTFTAverage = 0; ServerLoadSpec Load; };
struct ServerSpec { std::string Name, Class, Model; bool KeepAccelerator = false; int MinNumReplicas = 0, MaxBatchSize = 0; AllocationData CurrentAlloc, DesiredAlloc; };
struct OptimizerSpec { bool Unlimited = false, DelayedBestEffort = false; std::string SaturationPolicy; };
struct SystemSpec {
    std::vector<AcceleratorSpec> Accelerators; std::vector<ModelAcceleratorPerfData> Models;
    std::vector<ServiceClassSpec> ServiceClasses; std::vector<ServerSpec> Servers; OptimizerSpec Optimizer;
    std::vector<AcceleratorCount> Capacity;
};
struct AllocationSolution { std::map<std::string, AllocationData> Spec; };
const char* const DefaultServiceClassName = "Free";      // pkg/config/defaults.go:24
constexpr int DefaultServiceClassPriority = 100;           // :27-33
inline int SaturatedAllocationPolicyEnum(const std::string& s) {   // pkg/config/config.go:28-41
    if (s == "PriorityExhaustive") return WVA_POLICY_PRIORITY_EXHAUSTIVE;
    if (s == "PriorityRoundRobin") return WVA_POLICY_PRIORITY_ROUND_ROBIN;
    if (s == "RoundRobin") return WVA_POLICY_ROUND_ROBIN;
    return WVA_POLICY_NONE;
}
}  // namespace config

// One native context per process (wva_ctx_create is the slow call), like the Go shim keeps.
class NativeContext {
public:
    explicit NativeContext(int device = 0) {
        int rc = wva_ctx_create(device, &ctx_);
        if (rc != WVA_OK) throw Error(rc, wva_last_error(nullptr));
    }
    ~NativeContext() { wva_ctx_destroy(ctx_); }
    NativeContext(const NativeContext&) = delete;
    wva_ctx* get() const { return ctx_; }
    void check(int rc) const { if (rc != WVA_OK) throw Error(rc, wva_last_error(ctx_)); }
private:
    wva_ctx* ctx_ = nullptr;
};

namespace core {

class System;

// core.Allocation, pkg/core/allocation.go:13-24 (+ getters :209-252)
class Allocation {
public:
    const std::string& Accelerator() const { return accelerator; }
    int64_t NumReplicas() const { return numReplicas; }
    int64_t MaxBatchSize() const { return batchSize; }
    float MaxArrvRatePerReplica() const { return maxArrvRatePerReplica; }
    float MaxRPM() const { float t = maxArrvRatePerReplica * 1000.0f; return t * 60.0f; }   // :236-238
    float Cost() const { return cost; }
    float Value() const { return value; }
    float ITL() const { return itl; }
    float TTFT() const { return ttft; }
    float Rho() const { return rho; }
    bool Saturated(float totalRate) const { return totalRate > (float)numReplicas * MaxRPM(); }   // :254-256
    config::AllocationData AllocationData() const {                                               // :317-326
        config::AllocationData d; d.Accelerator = accelerator; d.NumReplicas = numReplicas; d.MaxBatch = batchSize;
        d.Cost = cost; d.ITLAverage = itl; d.TTFTAverage = ttft; return d;
    }
    std::string accelerator; int64_t numReplicas = 0, batchSize = 0;
    float cost = 0, value = 0, itl = 0, ttft = 0, rho = 0, maxArrvRatePerReplica = 0;
};

// core.Server, pkg/core/server.go
class Server {
public:
    const std::string& Name() const { return spec.Name; }
    const std::string& ServiceClassName() const { return serviceClassName; }
    const std::string& ModelName() const { return spec.Model; }
    bool KeepAccelerator() const { return spec.KeepAccelerator; }
    const config::ServerLoadSpec& Load() const { return spec.CurrentAlloc.Load; }
    int Priority() const { return priority; }
    // Server.Calculate (server.go:55-67): candidates for this server.  The first call after SetFromSpec
    // sizes EVERY server in one launch (the reference's AnalyzeModel works by side effect on the shared
    // System too, internal/modelanalyzer/analyzer.go:29-31); later calls read the cached result.
    void Calculate();
    const std::map<std::string, std::shared_ptr<core::Allocation>>& AllAllocations() const { return allAllocations; }
    const std::shared_ptr<core::Allocation>& Allocation() const { return allocation; }
    bool Saturated() const { return allocation && allocation->Saturated(spec.CurrentAlloc.Load.ArrivalRate); }
    config::ServerSpec spec;
    std::string serviceClassName;
    int priority = config::DefaultServiceClassPriority;
    int index = -1;
    System* system = nullptr;
    std::map<std::string, std::shared_ptr<core::Allocation>> allAllocations;
    std::shared_ptr<core::Allocation> allocation;      // allocated solution
};

struct AllocationByType { std::string name; int64_t count = 0, limit = 0; float cost = 0; };   // system.go:58-64

// core.System, pkg/core/system.go
class System {
public:
    explicit System(NativeContext& native) : native_(native) {}
    // H2D bytes of the last image upload or incremental update (wva_upload_bytes)
    int64_t UploadBytes() const { return wva_upload_bytes(native_.get()); }

    // SetFromSpec (system.go:82-90): interns names, resolves (class, model) targets, uploads the SoA image
    const config::OptimizerSpec& SetFromSpec(const config::SystemSpec& d) {
        spec_ = d;
        accNames_.clear(); typeNames_.clear(); modelNames_.clear(); servers_.clear(); serverOrder_.clear();
        std::map<std::string, int> accIdx, typeIdx, modelIdx;
        std::map<std::string, config::AcceleratorSpec> accs;
        for (const auto& a : d.Accelerators) { if (!accs.count(a.Name)) accNames_.push_back(a.Name); accs[a.Name] = a; }
        for (size_t i = 0; i < accNames_.size(); ++i) accIdx[accNames_[i]] = (int)i;
        for (const auto& n : accNames_) { const auto& t = accs[n].Type; if (!typeIdx.count(t)) { typeIdx[t] = (int)typeNames_.size(); typeNames_.push_back(t); } }
        capacity_.clear();
        for (const auto& c : d.Capacity) capacity_[c.Type] = c.Count;
        std::map<std::string, std::map<std::string, config::ModelAcceleratorPerfData>> models;
        for (const auto& pd : d.Models) { if (!models.count(pd.Name)) { modelIdx[pd.Name] = (int)modelNames_.size(); modelNames_.push_back(pd.Name); } models[pd.Name][pd.Acc] = pd; }
        std::map<std::string, std::pair<int, std::map<std::string, config::ModelTarget>>> classes;
        for (const auto& sc : d.ServiceClasses) {
            int pr = sc.Priority; if (pr < 1 || pr > 100) pr = config::DefaultServiceClassPriority;      // serviceclass.go:28-37
            auto& e = classes[sc.Name]; e.first = pr; e.second.clear();
            for (const auto& mt : sc.ModelTargets) e.second[mt.Model] = mt;
        }
        std::map<std::string, config::ServerSpec> srv;
        for (const auto& s : d.Servers) { if (!srv.count(s.Name)) serverOrder_.push_back(s.Name); srv[s.Name] = s; }

        const int S = (int)serverOrder_.size(), A = (int)accNames_.size(), M = (int)modelNames_.size(), T = (int)typeNames_.size();
        S_ = S; A_ = A; M_ = M; T_ = T;
        acc_cost.assign(A, 0); acc_mult.assign(A, 0); acc_type.assign(A, 0); type_cap.assign(T, 0);
        for (int a = 0; a < A; ++a) { const auto& s = accs[accNames_[a]]; acc_cost[a] = s.Cost; acc_mult[a] = s.Multiplicity; acc_type[a] = typeIdx[s.Type]; }
        for (int t = 0; t < T; ++t) { auto it = capacity_.find(typeNames_[t]); type_cap[t] = it == capacity_.end() ? 0 : it->second; }
        const size_t MA = (size_t)M * A;
        p_alpha.assign(MA, 0); p_beta.assign(MA, 0); p_gamma.assign(MA, 0); p_delta.assign(MA, 0);
        p_mb.assign(MA, 0); p_at.assign(MA, 0); p_cnt.assign(MA, 0); p_valid.assign(MA, 0);
        for (const auto& m : models) for (const auto& pa : m.second) {
            auto ai = accIdx.find(pa.first); if (ai == accIdx.end()) continue;
            size_t k = (size_t)modelIdx[m.first] * A + ai->second; const auto& pd = pa.second;
            p_alpha[k] = pd.Decode.Alpha; p_beta[k] = pd.Decode.Beta; p_gamma[k] = pd.Prefill.Gamma; p_delta[k] = pd.Prefill.Delta;
            p_mb[k] = pd.MaxBatchSize; p_at[k] = pd.AtTokens; p_cnt[k] = pd.AccCount; p_valid[k] = 1;
        }
        s_model.assign(S, -1); s_arr.assign(S, 0); s_in.assign(S, 0); s_out.assign(S, 0); s_ttft.assign(S, 0); s_itl.assign(S, 0);
        s_tps.assign(S, 0); s_tv.assign(S, 0); s_prio.assign(S, config::DefaultServiceClassPriority); s_minr.assign(S, 0);
        s_mb.assign(S, 0); s_keep.assign(S, 0); s_cacc.assign(S, WVA_ACC_NONE); s_crep.assign(S, 0); s_ccost.assign(S, 0);
        accIdx_ = accIdx; modelIdx_ = modelIdx; classes_ = classes;
        for (int i = 0; i < S; ++i) {
            auto sv = std::make_shared<Server>();
            sv->spec = srv[serverOrder_[i]]; sv->index = i; sv->system = this;
            fillServerRow(*sv, i);
            servers_[serverOrder_[i]] = sv;
        }
        uploadImage(s_keep);
        analyzed_ = false; created_ = false;
        return spec_.Optimizer;
    }

    // Incremental updates (system.go:98-171) go into the RESIDENT image: a server's row (54 bytes) or the T capacity
    // counters cross PCIe, not the image (wva_system_update_servers / wva_system_remove_server / wva_system_set_capacity).
    // "replace if it already exists" and the not-found errors (returned as false) are the reference's.  Accelerator
    // changes alter the table layout (A) and rebuild the image.
    void AddServerFromSpec(const config::ServerSpec& s) {
        bool found = false;
        for (auto& e : spec_.Servers) if (e.Name == s.Name) { e = s; found = true; }
        if (!found) spec_.Servers.push_back(s);
        auto it = servers_.find(s.Name);
        const int idx = it != servers_.end() ? it->second->index : S_;
        if (idx == S_) {                                   // append: grow the host arrays by one row
            serverOrder_.push_back(s.Name);
            s_model.push_back(-1); s_arr.push_back(0); s_in.push_back(0); s_out.push_back(0); s_ttft.push_back(0); s_itl.push_back(0); s_tps.push_back(0);
            s_tv.push_back(0); s_prio.push_back(config::DefaultServiceClassPriority); s_minr.push_back(0); s_mb.push_back(0); s_keep.push_back(0);
            s_cacc.push_back(WVA_ACC_NONE); s_crep.push_back(0); s_ccost.push_back(0);
            ++S_;
        }
        auto sv = std::make_shared<Server>();
        sv->spec = s; sv->index = idx; sv->system = this;
        fillServerRow(*sv, idx);
        servers_[s.Name] = sv;
        wva_system_soa row = serverRows(idx, 1);
        int rc = wva_system_update_servers(native_.get(), idx, 1, &row);
        if (rc == WVA_ECAPACITY) uploadImage(s_keep);      // no spare row left: the whole image goes up again
        else native_.check(rc);
        analyzed_ = false; created_ = false;
    }
    bool RemoveServer(const std::string& name) {
        auto it = servers_.find(name);
        if (it == servers_.end()) return false;
        for (size_t i = spec_.Servers.size(); i-- > 0;) if (spec_.Servers[i].Name == name) spec_.Servers.erase(spec_.Servers.begin() + (long)i);
        const int idx = it->second->index, last = S_ - 1;
        servers_.erase(it);
        if (idx != last) {                                 // the last server moves into the freed slot (host arrays in step with the device)
            const std::string moved = serverOrder_[(size_t)last];
            serverOrder_[(size_t)idx] = moved; servers_[moved]->index = idx;
            s_model[idx] = s_model[last]; s_arr[idx] = s_arr[last]; s_in[idx] = s_in[last]; s_out[idx] = s_out[last]; s_ttft[idx] = s_ttft[last];
            s_itl[idx] = s_itl[last]; s_tps[idx] = s_tps[last]; s_tv[idx] = s_tv[last]; s_prio[idx] = s_prio[last]; s_minr[idx] = s_minr[last];
            s_mb[idx] = s_mb[last]; s_keep[idx] = s_keep[last]; s_cacc[idx] = s_cacc[last]; s_crep[idx] = s_crep[last]; s_ccost[idx] = s_ccost[last];
        }
        serverOrder_.pop_back();
        s_model.pop_back(); s_arr.pop_back(); s_in.pop_back(); s_out.pop_back(); s_ttft.pop_back(); s_itl.pop_back(); s_tps.pop_back(); s_tv.pop_back();
        s_prio.pop_back(); s_minr.pop_back(); s_mb.pop_back(); s_keep.pop_back(); s_cacc.pop_back(); s_crep.pop_back(); s_ccost.pop_back();
        --S_;
        native_.check(wva_system_remove_server(native_.get(), idx));
        analyzed_ = false; created_ = false;
        return true;
    }
    void AddAcceleratorFromSpec(const config::AcceleratorSpec& a) {
        config::SystemSpec d = spec_;
        bool found = false;
        for (auto& e : d.Accelerators) if (e.Name == a.Name) { e = a; found = true; }
        if (!found) d.Accelerators.push_back(a);
        SetFromSpec(d);
    }
    bool RemoveAccelerator(const std::string& name) {
        config::SystemSpec d = spec_;
        const size_t before = d.Accelerators.size();
        for (size_t i = d.Accelerators.size(); i-- > 0;) if (d.Accelerators[i].Name == name) d.Accelerators.erase(d.Accelerators.begin() + (long)i);
        if (d.Accelerators.size() == before) return false;
        SetFromSpec(d);
        return true;
    }
    void SetCountFromSpec(const config::AcceleratorCount& c) {
        bool found = false;
        for (auto& e : spec_.Capacity) if (e.Type == c.Type) { e = c; found = true; }
        if (!found) spec_.Capacity.push_back(c);
        capacity_[c.Type] = c.Count;
        for (int t = 0; t < T_; ++t) if (typeNames_[t] == c.Type) type_cap[t] = c.Count;   // a type no accelerator has never matters
        native_.check(wva_system_set_capacity(native_.get(), type_cap.data()));
    }
    void SetCapacityFromSpec(const std::vector<config::AcceleratorCount>& v) { for (const auto& c : v) SetCountFromSpec(c); }

    // Allocation.Scale (allocation.go:165-188): a fresh CreateAllocation on the allocation's own accelerator
    // and the change in replicas.  {nullptr, 0} where the reference returns nil (and where it would
    // dereference the nil result of CreateAllocation).
    std::pair<std::shared_ptr<Allocation>, int64_t> Scale(const Allocation& a, const std::string& serverName) {
        auto sv = GetServer(serverName);
        if (!sv) return {nullptr, 0};
        int ai = -1;
        for (int i = 0; i < A_; ++i) if (accNames_[i] == a.accelerator) ai = i;
        if (ai < 0) return {nullptr, 0};
        createAll();
        auto al = created_alloc_[(size_t)sv->index * A_ + ai];
        if (!al) return {nullptr, 0};
        return {std::make_shared<Allocation>(*al), al->numReplicas - a.numReplicas};
    }

    // Allocation.ReAllocate (allocation.go:190-207): the cheapest CreateAllocation over ALL accelerators
    // (not only the server's candidates); map order canonicalised to ascending accelerator index.
    std::pair<std::shared_ptr<Allocation>, std::string> ReAllocate(const std::string& serverName) {
        auto sv = GetServer(serverName);
        if (!sv) return {nullptr, std::string()};
        createAll();
        float minVal = 0.0f;
        std::shared_ptr<Allocation> best;
        for (int a = 0; a < A_; ++a) {
            auto al = created_alloc_[(size_t)sv->index * A_ + a];
            if (!al) continue;
            if (minVal == 0.0f || al->value < minVal) { minVal = al->value; best = al; }
        }
        if (!best) return {nullptr, std::string()};
        return {std::make_shared<Allocation>(*best), best->accelerator};
    }

    const std::map<std::string, std::shared_ptr<Server>>& Servers() const { return servers_; }
    std::shared_ptr<Server> GetServer(const std::string& n) const { auto it = servers_.find(n); return it == servers_.end() ? nullptr : it->second; }
    const std::vector<std::string>& AcceleratorNames() const { return accNames_; }
    const std::map<std::string, int64_t>& Capacities() const { return capacity_; }

    // System.Calculate (system.go:262-272): every server
    void Calculate() { analyzeAll(); }

    // one launch for all pairs; fills Server.allAllocations (server.go:57-66)
    void analyzeAll() {
        if (analyzed_) return;
        const size_t n = (size_t)S_ * A_;
        std::vector<int32_t> acc(n); std::vector<int64_t> rep(n), bat(n); std::vector<float> cost(n), val(n), itl(n), ttft(n), rho(n), arrv(n);
        std::vector<uint8_t> fe(n);
        wva_alloc_soa o{acc.data(), rep.data(), bat.data(), cost.data(), val.data(), itl.data(), ttft.data(), rho.data(), arrv.data()};
        if (n) native_.check(wva_analyze_pairs(native_.get(), &o, fe.data()));
        for (int s = 0; s < S_; ++s) {
            auto& sv = servers_[serverOrder_[s]];
            sv->allAllocations.clear(); sv->allocation.reset();
            for (int a = 0; a < A_; ++a) {
                size_t i = (size_t)s * A_ + a; if (!fe[i]) continue;
                auto al = std::make_shared<Allocation>();
                al->accelerator = acc[i] >= 0 ? accNames_[acc[i]] : std::string();
                al->numReplicas = rep[i]; al->batchSize = bat[i]; al->cost = cost[i]; al->value = val[i]; al->itl = itl[i];
                al->ttft = ttft[i]; al->rho = rho[i]; al->maxArrvRatePerReplica = arrv[i];
                sv->allAllocations[accNames_[a]] = al;
            }
        }
        analyzed_ = true;
    }

    // CreateAllocation(server, accelerator) for every pair, candidate or not (allocation.go:27-163): the pair kernel
    // honours Server.GetCandidateAccelerators, so the image goes to a SECOND context with keepAccelerator cleared and
    // every pair is sized there in one launch -- the main context keeps its candidates, assignment and totals, so
    // Solve() -> ReAllocate()/Scale() -> AllocateByType() works as in the reference.  value = cost as
    // CreateAllocation leaves it (:161) -- the transition penalty belongs to Server.Calculate.
    void createAll() {
        if (created_) return;
        const size_t n = (size_t)S_ * A_;
        created_alloc_.assign(n, nullptr);
        if (n) {
            if (!scratch_) scratch_.reset(new NativeContext(0));
            std::vector<uint8_t> none((size_t)S_, 0);
            wva_system_soa h = imageView(none);
            scratch_->check(wva_system_upload(scratch_->get(), &h));
            std::vector<int32_t> acc(n); std::vector<int64_t> rep(n), bat(n); std::vector<float> cost(n), val(n), itl(n), ttft(n), rho(n), arrv(n);
            std::vector<uint8_t> fe(n);
            wva_alloc_soa o{acc.data(), rep.data(), bat.data(), cost.data(), val.data(), itl.data(), ttft.data(), rho.data(), arrv.data()};
            scratch_->check(wva_analyze_pairs(scratch_->get(), &o, fe.data()));
            for (size_t i = 0; i < n; ++i) {
                if (!fe[i]) continue;
                auto al = std::make_shared<Allocation>();
                al->accelerator = acc[i] >= 0 ? accNames_[acc[i]] : std::string();
                al->numReplicas = rep[i]; al->batchSize = bat[i]; al->cost = cost[i]; al->value = cost[i]; al->itl = itl[i];
                al->ttft = ttft[i]; al->rho = rho[i]; al->maxArrvRatePerReplica = arrv[i];
                created_alloc_[i] = al;
            }
        }
        created_ = true;
    }

    // Solver.Solve dispatch used by solver::Solver
    void solve(const config::OptimizerSpec& os) {
        analyzeAll();
        wva_optimizer_spec sp{os.Unlimited ? 1 : 0, os.DelayedBestEffort ? 1 : 0, config::SaturatedAllocationPolicyEnum(os.SaturationPolicy)};
        std::vector<int32_t> key(S_), acc(S_); std::vector<int64_t> rep(S_), bat(S_); std::vector<float> cost(S_), val(S_), itl(S_), ttft(S_), rho(S_), arrv(S_);
        wva_alloc_soa o{acc.data(), rep.data(), bat.data(), cost.data(), val.data(), itl.data(), ttft.data(), rho.data(), arrv.data()};
        native_.check(wva_solve(native_.get(), &sp, key.data(), &o));
        for (int s = 0; s < S_; ++s) {
            auto& sv = servers_[serverOrder_[s]];
            sv->allocation.reset();
            if (key[s] < 0) continue;
            // the chosen *Allocation is shared with allAllocations in the reference; best effort may have scaled it
            auto al = sv->allAllocations[accNames_[key[s]]];
            if (!al) al = std::make_shared<Allocation>();
            al->accelerator = acc[s] >= 0 ? accNames_[acc[s]] : std::string();
            al->numReplicas = rep[s]; al->batchSize = bat[s]; al->cost = cost[s]; al->value = val[s]; al->itl = itl[s];
            al->ttft = ttft[s]; al->rho = rho[s]; al->maxArrvRatePerReplica = arrv[s];
            sv->allocation = al;
            sv->spec.DesiredAlloc = al->AllocationData(); sv->spec.DesiredAlloc.Load = sv->spec.CurrentAlloc.Load;   // UpdateDesiredAlloc :148-155
        }
        if (!os.Unlimited) analyzed_ = false;   // candidates were mutated in place on the device (greedy.go:208-212)
    }

    // AllocateByType, system.go:271-300
    void AllocateByType() {
        std::vector<int64_t> count(T_); std::vector<float> cost(T_);
        native_.check(wva_allocate_by_type(native_.get(), count.data(), cost.data()));
        allocationByType_.clear();
        // only types that received an allocation appear in the reference's map
        std::vector<char> used(T_, 0);
        for (int s = 0; s < S_; ++s) { auto& sv = servers_[serverOrder_[s]]; if (!sv->allocation || sv->allocation->accelerator.empty() || s_model[s] < 0) continue;
            for (int a = 0; a < A_; ++a) if (accNames_[a] == sv->allocation->accelerator) used[acc_type[a]] = 1; }
        for (int t = 0; t < T_; ++t) if (used[t]) { AllocationByType e; e.name = typeNames_[t]; e.count = count[t]; e.cost = cost[t];
            auto it = capacity_.find(e.name); e.limit = it == capacity_.end() ? 0 : it->second; allocationByType_[e.name] = e; }
    }
    const std::map<std::string, AllocationByType>& AllocationByTypeMap() const { return allocationByType_; }

    // GenerateSolution, system.go:303-319
    config::AllocationSolution GenerateSolution() const {
        config::AllocationSolution sol;
        for (const auto& kv : servers_) { if (!kv.second->allocation) continue; auto d = kv.second->allocation->AllocationData(); d.Load = kv.second->Load(); sol.Spec[kv.first] = d; }
        return sol;
    }
    int64_t SolutionTimeUsec() const { return wva_solution_time_usec(native_.get()); }
    NativeContext& native() { return native_; }

private:
    wva_system_soa imageView(const std::vector<uint8_t>& keep) const {
        wva_system_soa h{};
        h.n_servers = S_; h.n_accels = A_; h.n_models = M_; h.n_types = T_;
        h.acc_cost = acc_cost.data(); h.acc_multiplicity = acc_mult.data(); h.acc_type = acc_type.data(); h.type_capacity = type_cap.data();
        h.perf_alpha = p_alpha.data(); h.perf_beta = p_beta.data(); h.perf_gamma = p_gamma.data(); h.perf_delta = p_delta.data();
        h.perf_max_batch = p_mb.data(); h.perf_at_tokens = p_at.data(); h.perf_acc_count = p_cnt.data(); h.perf_valid = p_valid.data();
        h.srv_model = s_model.data(); h.srv_arrival_rpm = s_arr.data(); h.srv_in_tokens = s_in.data(); h.srv_out_tokens = s_out.data();
        h.srv_slo_ttft = s_ttft.data(); h.srv_slo_itl = s_itl.data(); h.srv_slo_tps = s_tps.data(); h.srv_target_valid = s_tv.data();
        h.srv_priority = s_prio.data(); h.srv_min_replicas = s_minr.data(); h.srv_max_batch = s_mb.data(); h.srv_keep_acc = keep.data();
        h.srv_cur_acc = s_cacc.data(); h.srv_cur_replicas = s_crep.data(); h.srv_cur_cost = s_ccost.data();
        return h;
    }
    // rows [first, first+count) of the server arrays as a wva_system_soa for wva_system_update_servers
    wva_system_soa serverRows(int first, int count) const {
        wva_system_soa h = imageView(s_keep);
        h.n_servers = count;
        h.srv_model += first; h.srv_arrival_rpm += first; h.srv_in_tokens += first; h.srv_out_tokens += first; h.srv_slo_ttft += first;
        h.srv_slo_itl += first; h.srv_slo_tps += first; h.srv_target_valid += first; h.srv_priority += first; h.srv_min_replicas += first;
        h.srv_max_batch += first; h.srv_keep_acc += first; h.srv_cur_acc += first; h.srv_cur_replicas += first; h.srv_cur_cost += first;
        return h;
    }
    void uploadImage(const std::vector<uint8_t>& keep) {
        wva_system_soa h = imageView(keep);
        native_.check(wva_system_upload(native_.get(), &h));
    }
    // one server's row of the SoA image: model index, (class, model) target lookup, priority, current allocation
    void fillServerRow(Server& sv, int i) {
        sv.serviceClassName = sv.spec.Class.empty() ? config::DefaultServiceClassName : sv.spec.Class;   // server.go:36-39
        sv.priority = config::DefaultServiceClassPriority;
        auto mi = modelIdx_.find(sv.spec.Model); s_model[i] = mi == modelIdx_.end() ? -1 : mi->second;
        const auto& ld = sv.spec.CurrentAlloc.Load;
        s_arr[i] = ld.ArrivalRate; s_in[i] = ld.AvgInTokens; s_out[i] = ld.AvgOutTokens;
        s_tv[i] = 0; s_itl[i] = 0; s_ttft[i] = 0; s_tps[i] = 0; s_prio[i] = config::DefaultServiceClassPriority;
        auto ci = classes_.find(sv.serviceClassName);
        if (ci != classes_.end()) {
            sv.priority = ci->second.first; s_prio[i] = sv.priority;
            auto ti = ci->second.second.find(sv.spec.Model);
            if (ti != ci->second.second.end()) { s_tv[i] = 1; s_itl[i] = ti->second.SLO_ITL; s_ttft[i] = ti->second.SLO_TTFT; s_tps[i] = ti->second.SLO_TPS; }
        }
        s_minr[i] = sv.spec.MinNumReplicas; s_mb[i] = sv.spec.MaxBatchSize; s_keep[i] = sv.spec.KeepAccelerator ? 1 : 0;
        const auto& ca = sv.spec.CurrentAlloc.Accelerator;
        if (ca.empty()) s_cacc[i] = WVA_ACC_NONE; else { auto ai = accIdx_.find(ca); s_cacc[i] = ai == accIdx_.end() ? WVA_ACC_UNKNOWN : ai->second; }
        s_crep[i] = (int32_t)sv.spec.CurrentAlloc.NumReplicas; s_ccost[i] = sv.spec.CurrentAlloc.Cost;
    }
    std::unique_ptr<NativeContext> scratch_;
    std::map<std::string, int> accIdx_, modelIdx_;
    std::map<std::string, std::pair<int, std::map<std::string, config::ModelTarget>>> classes_;
    NativeContext& native_;
    config::SystemSpec spec_;
    std::vector<std::string> accNames_, typeNames_, modelNames_, serverOrder_;
    std::map<std::string, std::shared_ptr<Server>> servers_;
    std::map<std::string, int64_t> capacity_;
    std::map<std::string, AllocationByType> allocationByType_;
    int S_ = 0, A_ = 0, M_ = 0, T_ = 0;
    bool analyzed_ = false, created_ = false;
    std::vector<std::shared_ptr<Allocation>> created_alloc_;
    std::vector<float> acc_cost, p_alpha, p_beta, p_gamma, p_delta, s_arr, s_ttft, s_itl, s_tps, s_ccost;
    std::vector<int32_t> acc_mult, acc_type, p_mb, p_at, p_cnt, s_model, s_in, s_out, s_prio, s_minr, s_mb, s_cacc, s_crep;
    std::vector<int64_t> type_cap;
    std::vector<uint8_t> p_valid, s_tv, s_keep;
};

inline void Server::Calculate() { system->analyzeAll(); }

}  // namespace core

namespace solver {
// solver.Solver, pkg/solver/solver.go:13-60
class Solver {
public:
    Solver(const config::OptimizerSpec& spec, core::System& system) : spec_(spec), system_(system) {}
    void Solve() { system_.solve(spec_); }
private:
    config::OptimizerSpec spec_; core::System& system_;
};
// solver.Optimizer, pkg/solver/optimizer.go:11-38
class Optimizer {
public:
    explicit Optimizer(const config::OptimizerSpec& spec) : spec_(spec) {}      // NewOptimizerFromSpec
    void Optimize(core::System& system) { Solver s(spec_, system); s.Solve(); solutionTimeMsec_ = system.SolutionTimeUsec() / 1000; }
    int64_t SolutionTimeMsec() const { return solutionTimeMsec_; }
private:
    config::OptimizerSpec spec_; int64_t solutionTimeMsec_ = 0;
};
}  // namespace solver

namespace manager {
// manager.Manager, pkg/manager/manager.go:8-27
class Manager {
public:
    Manager(core::System& system, solver::Optimizer& optimizer) : system_(system), optimizer_(optimizer) {}
    void Optimize() { optimizer_.Optimize(system_); system_.AllocateByType(); }
private:
    core::System& system_; solver::Optimizer& optimizer_;
};
}  // namespace manager

namespace modelanalyzer {
// interfaces.ModelAcceleratorAllocation / ModelAnalyzeResponse, internal/interfaces/types.go:6-20
struct ModelAcceleratorAllocation { std::shared_ptr<core::Allocation> Allocation; double RequiredPrefillQPS = 0, RequiredDecodeQPS = 0; std::string Reason; };
struct ModelAnalyzeResponse { std::map<std::string, ModelAcceleratorAllocation> Allocations; };
// ModelAnalyzer, internal/modelanalyzer/analyzer.go:13-34 (serverName = utils.FullName(va.Name, va.Namespace))
class ModelAnalyzer {
public:
    explicit ModelAnalyzer(core::System& system) : system_(system) {}
    ModelAnalyzeResponse AnalyzeModel(const std::string& vaName, const std::string& vaNamespace) {
        ModelAnalyzeResponse r;
        auto sv = system_.GetServer(vaName + ":" + vaNamespace);
        if (!sv) return r;
        sv->Calculate();
        for (const auto& kv : sv->AllAllocations()) {            // CreateModelAnalyzeResponseFromAllocations, utils.go:9-24
            ModelAcceleratorAllocation m; m.Allocation = kv.second;
            m.RequiredPrefillQPS = (double)(kv.second->MaxArrvRatePerReplica() * 1000.0f);
            m.RequiredDecodeQPS = m.RequiredPrefillQPS; m.Reason = "markovian analysis";
            r.Allocations[kv.first] = m;
        }
        return r;
    }
private:
    core::System& system_;
};
}  // namespace modelanalyzer

namespace optimizer {
struct OptimizedAlloc { std::string Accelerator; int64_t NumReplicas = 0; };   // api/v1alpha1 OptimizedAlloc (LastRunTime set by caller)
struct VariantRef { std::string Name, Namespace; };
// VariantAutoscalingsEngine, internal/optimizer/optimizer.go:16-54
class VariantAutoscalingsEngine {
public:
    VariantAutoscalingsEngine(manager::Manager& m, core::System& s) : manager_(m), system_(s) {}
    // returns the map keyed by bare va.Name; throws Error(WVA_ENOSOLUTION) when the solution is empty (:38-40)
    std::map<std::string, OptimizedAlloc> Optimize(const std::vector<VariantRef>& vaList) {
        manager_.Optimize();
        auto sol = system_.GenerateSolution();
        if (sol.Spec.empty()) throw Error(WVA_ENOSOLUTION, "no feasible allocations found for all variants: ");
        std::map<std::string, OptimizedAlloc> out;
        for (const auto& va : vaList) {
            auto it = sol.Spec.find(va.Name + ":" + va.Namespace);      // utils.CreateOptimizedAlloc, utils.go:314-331
            if (it == sol.Spec.end()) continue;
            out[va.Name] = OptimizedAlloc{it->second.Accelerator, it->second.NumReplicas};
        }
        return out;
    }
private:
    manager::Manager& manager_; core::System& system_;
};
}  // namespace optimizer

}  // namespace wva
