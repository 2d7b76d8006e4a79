// wva_adapters.hpp — the data formats either side of the path, in the C++ host mirror:
//
//   IN : the adapters of internal/utils/utils.go:108-311 — CreateSystemData (accelerator + service-class ConfigMaps),
//        AddModelAcceleratorProfileToSystemData (the VariantAutoscaling's per-accelerator perf strings),
//        AddServerInfoToSystemData (current allocation + load strings of the CR status) — producing the
//        config::SystemSpec that core::System::SetFromSpec turns into the SoA image in one pass (names interned to
//        dense indices, (class, model) targets resolved per server: wva_host.hpp).
//   OUT: config.AllocationSolution as JSON with the reference's field names (pkg/config/types.go:123-143), the
//        OptimizedAlloc the controller writes back (utils.CreateOptimizedAlloc, utils.go:314-331), and — what the
//        reference computes but drops (api/v1alpha1/variantautoscaling_types.go:138-149 carries accelerator + replicas
//        only) — the chosen batch size, ITL/TTFT/rho and the candidate sweep's winner per server.
//
// String parsing follows strconv.ParseFloat(s, 32): the NEAREST float32 of the decimal string (std::strtof, one
// rounding — not a double rounded twice), NaN / Inf / malformed -> the reference's fallback (0, or skip the entry).
#pragma once

#include <cctype>
#include <cerrno>
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "wva_host.hpp"

namespace wva {
namespace utils {

// strconv.ParseFloat(s, 32): ok = false on syntax error / empty / trailing garbage / out of range
inline bool ParseFloat32(const std::string& s, float* out) {
    if (s.empty() || std::isspace((unsigned char)s[0])) return false;      // strtof would skip leading blanks, Go does not
    const char* b = s.c_str();
    char* e = nullptr;
    errno = 0;
    const float v = std::strtof(b, &e);
    if (e == b || *e != '\0') return false;
    if (errno == ERANGE && std::isinf(v)) return false;          // Go: value out of range is an error (returns +-Inf, err)
    *out = v;
    return true;
}
// utils.CheckValue, utils.go:338-340
inline bool CheckValue(double x) { return !(std::isnan(x) || std::isinf(x)); }
// utils.FullName, utils.go:333-336
inline std::string FullName(const std::string& name, const std::string& ns) { return name + ":" + ns; }

// interfaces.ServiceClass (what yaml.Unmarshal of one service-class ConfigMap entry yields; the YAML decoding itself is
// a library call in the reference, not adapter logic)
struct ServiceClassEntry { std::string Model; double SLOTPOT = 0, SLOTTFT = 0; };
struct ServiceClass { std::string Name; int Priority = 0; std::vector<ServiceClassEntry> Data; };

// llmdVariantAutoscalingV1alpha1.AcceleratorProfile (api/v1alpha1): perf parameters arrive as strings
struct AcceleratorProfile {
    std::string Acc; int AccCount = 0, MaxBatchSize = 0;
    std::map<std::string, std::string> DecodeParms, PrefillParms;       // "alpha","beta" / "gamma","delta"
};
// the fields of a VariantAutoscaling that AddServerInfoToSystemData reads
struct VariantAutoscaling {
    std::string Name, Namespace, ModelID, AcceleratorLabel;               // label inference.optimization/acceleratorName
    std::vector<AcceleratorProfile> Accelerators;                         // Spec.ModelProfile.Accelerators
    struct { std::string Accelerator; int NumReplicas = 0, MaxBatch = 0; std::string VariantCost, ITLAverage, TTFTAverage;
             struct { std::string ArrivalRate, AvgInputTokens, AvgOutputTokens; } Load; } CurrentAlloc;   // Status.CurrentAlloc
};

// CreateSystemData, utils.go:108-184.  acceleratorCm: name -> {"device": type, "cost": "32.00"}.
// Entries whose cost does not parse are skipped (warning in the reference); WVA runs the optimizer in unlimited mode.
inline config::SystemSpec CreateSystemData(const std::map<std::string, std::map<std::string, std::string>>& acceleratorCm,
                                           const std::vector<ServiceClass>& serviceClasses) {
    config::SystemSpec sd;
    for (const auto& kv : acceleratorCm) {
        float cost = 0;
        auto ci = kv.second.find("cost");
        if (ci == kv.second.end() || !ParseFloat32(ci->second, &cost)) continue;
        config::AcceleratorSpec a;
        a.Name = kv.first;
        auto di = kv.second.find("device");
        a.Type = di == kv.second.end() ? std::string() : di->second;
        a.Multiplicity = 1;                                               // utils.go:131
        a.Cost = cost;
        sd.Accelerators.push_back(a);
    }
    for (const auto& sc : serviceClasses) {
        config::ServiceClassSpec spec;
        spec.Name = sc.Name; spec.Priority = sc.Priority;
        for (const auto& e : sc.Data) spec.ModelTargets.push_back({e.Model, (float)e.SLOTPOT, (float)e.SLOTTFT, 0.0f});   // :158-164
        sd.ServiceClasses.push_back(spec);
    }
    sd.Optimizer.Unlimited = true;                                        // :170-173
    return sd;
}

// AddModelAcceleratorProfileToSystemData, utils.go:187-237.  Returns "" or the reference's error text.
inline std::string AddModelAcceleratorProfileToSystemData(config::SystemSpec& sd, const std::string& modelName, const AcceleratorProfile& p) {
    if (p.DecodeParms.size() < 2) return "length of decodeParms should be 2";
    float alpha = 0, beta = 0, gamma = 0, delta = 0;
    auto get = [](const std::map<std::string, std::string>& m, const char* k, float* v) {
        auto it = m.find(k);
        return ParseFloat32(it == m.end() ? std::string() : it->second, v);
    };
    if (!get(p.DecodeParms, "alpha", &alpha)) return "strconv.ParseFloat: parsing alpha: invalid syntax";
    if (!get(p.DecodeParms, "beta", &beta)) return "strconv.ParseFloat: parsing beta: invalid syntax";
    if (p.PrefillParms.size() < 2) return "length of prefillParms should be 2";
    if (!get(p.PrefillParms, "gamma", &gamma)) return "strconv.ParseFloat: parsing gamma: invalid syntax";
    if (!get(p.PrefillParms, "delta", &delta)) return "strconv.ParseFloat: parsing delta: invalid syntax";
    config::ModelAcceleratorPerfData pd;
    pd.Name = modelName; pd.Acc = p.Acc; pd.AccCount = p.AccCount; pd.MaxBatchSize = p.MaxBatchSize;   // AtTokens stays 0 (:222-235)
    pd.Decode.Alpha = alpha; pd.Decode.Beta = beta; pd.Prefill.Gamma = gamma; pd.Prefill.Delta = delta;
    sd.Models.push_back(pd);
    return std::string();
}

// AddServerInfoToSystemData, utils.go:240-311.  scaleToZero = (os.Getenv("WVA_SCALE_TO_ZERO") == "true").
inline void AddServerInfoToSystemData(config::SystemSpec& sd, const VariantAutoscaling& va, const std::string& className, bool scaleToZero) {
    auto num = [](const std::string& s) { float v = 0; if (!ParseFloat32(s, &v) || !CheckValue((double)v)) v = 0; return v; };   // :247-255, :264-272
    config::ServerSpec sv;
    sv.Name = FullName(va.Name, va.Namespace);
    sv.Class = className;
    sv.Model = va.ModelID;
    sv.KeepAccelerator = true;                                            // :291
    sv.MinNumReplicas = scaleToZero ? 0 : 1;                              // :281-284
    sv.CurrentAlloc.Accelerator = va.CurrentAlloc.Accelerator;
    sv.CurrentAlloc.NumReplicas = va.CurrentAlloc.NumReplicas;
    sv.CurrentAlloc.MaxBatch = va.CurrentAlloc.MaxBatch;
    sv.CurrentAlloc.Cost = num(va.CurrentAlloc.VariantCost);
    sv.CurrentAlloc.ITLAverage = num(va.CurrentAlloc.ITLAverage);
    sv.CurrentAlloc.TTFTAverage = num(va.CurrentAlloc.TTFTAverage);
    sv.CurrentAlloc.Load.ArrivalRate = num(va.CurrentAlloc.Load.ArrivalRate);
    sv.CurrentAlloc.Load.AvgInTokens = (int)num(va.CurrentAlloc.Load.AvgInputTokens);     // int(float64) truncation, :258-259
    sv.CurrentAlloc.Load.AvgOutTokens = (int)num(va.CurrentAlloc.Load.AvgOutputTokens);
    for (const auto& ap : va.Accelerators)                                // :297-306: batch override of the labelled accelerator
        if (ap.Acc == va.AcceleratorLabel) { if (ap.MaxBatchSize > 0) sv.MaxBatchSize = ap.MaxBatchSize; break; }
    sd.Servers.push_back(sv);
}

// ---- OUT -----------------------------------------------------------------------------------------------------------

// float32 the way Go's encoding/json writes it: shortest digits that round-trip as float32, 'e' form only for
// exponents < -6 or >= 21 (encoding/json floatEncoder), "e-07" style two-digit negative exponents cleaned to "e-7".
inline std::string JsonFloat32(float f) {
    if (std::isnan(f) || std::isinf(f)) return "null";                    // json.Marshal fails on these; callers check CheckValue first
    if (f == 0.0f) return std::signbit(f) ? "-0" : "0";
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, f, std::chars_format::scientific);
    std::string sci(buf, r.ptr);                                          // d[.ddd]e[+-]XX, shortest round-trip digits
    const size_t epos = sci.find('e');
    std::string mant = sci.substr(0, epos);
    const int exp10 = std::atoi(sci.c_str() + epos + 1);
    const bool neg = mant[0] == '-';
    if (neg) mant.erase(0, 1);
    std::string digits;
    for (char c : mant) if (c != '.') digits.push_back(c);
    std::string out = neg ? "-" : "";
    if (exp10 < -6 || exp10 >= 21) {
        out += digits.substr(0, 1);
        if (digits.size() > 1) out += "." + digits.substr(1);
        out += (exp10 < 0 ? "e-" : "e+") + std::to_string(exp10 < 0 ? -exp10 : exp10);
        return out;
    }
    if (exp10 < 0) { out += "0." + std::string((size_t)(-exp10 - 1), '0') + digits; return out; }
    if ((int)digits.size() <= exp10 + 1) { out += digits + std::string((size_t)(exp10 + 1 - (int)digits.size()), '0'); return out; }
    out += digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
    return out;
}
inline std::string JsonString(const std::string& s) {
    std::string o = "\"";
    for (unsigned char c : s) {
        if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back((char)c); }
        else if (c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o.push_back((char)c);
    }
    return o + "\"";
}
// config.AllocationData / AllocationSolution, pkg/config/types.go:123-143 (field order and json tags of the struct;
// encoding/json writes map keys sorted)
inline std::string ToJSON(const config::AllocationData& d) {
    std::ostringstream o;
    o << "{\"accelerator\":" << JsonString(d.Accelerator) << ",\"numReplicas\":" << d.NumReplicas << ",\"maxBatch\":" << d.MaxBatch
      << ",\"cost\":" << JsonFloat32(d.Cost) << ",\"itlAverage\":" << JsonFloat32(d.ITLAverage) << ",\"ttftAverage\":" << JsonFloat32(d.TTFTAverage)
      << ",\"load\":{\"arrivalRate\":" << JsonFloat32(d.Load.ArrivalRate) << ",\"avgInTokens\":" << d.Load.AvgInTokens
      << ",\"avgOutTokens\":" << d.Load.AvgOutTokens << "}}";
    return o.str();
}
inline std::string ToJSON(const config::AllocationSolution& sol) {
    std::ostringstream o;
    o << "{\"allocations\":{";
    bool first = true;
    for (const auto& kv : sol.Spec) { o << (first ? "" : ",") << JsonString(kv.first) << ":" << ToJSON(kv.second); first = false; }
    o << "}}";
    return o.str();
}

// utils.CreateOptimizedAlloc (utils.go:314-331) plus the outputs the reference computes and drops: the batch size and the
// predicted ITL / TTFT / utilisation of the chosen allocation, and the candidate sweep's winner for the server.
struct OptimizedAllocExt {
    std::string Accelerator; int64_t NumReplicas = 0;                     // what OptimizedAlloc carries
    int64_t MaxBatch = 0; float Cost = 0, ITLAverage = 0, TTFTAverage = 0, Rho = 0, MaxArrvRatePerReplica = 0;
    bool HaveSweep = false;
    std::string SweepAccelerator; int SweepReplicas = 0, SweepBatch = 0; float SweepCost = 0, SweepITL = 0, SweepTTFT = 0, SweepRho = 0;
};
inline bool CreateOptimizedAllocExt(core::System& system, const std::string& name, const std::string& ns, const wva_grid_best* sweep,
                                    OptimizedAllocExt* out) {
    auto sv = system.GetServer(FullName(name, ns));
    if (!sv || !sv->Allocation()) return false;                           // "server %s not found"
    const auto& al = *sv->Allocation();
    out->Accelerator = al.Accelerator(); out->NumReplicas = al.NumReplicas(); out->MaxBatch = al.MaxBatchSize();
    out->Cost = al.Cost(); out->ITLAverage = al.ITL(); out->TTFTAverage = al.TTFT(); out->Rho = al.Rho();
    out->MaxArrvRatePerReplica = al.MaxArrvRatePerReplica();
    out->HaveSweep = false;
    if (sweep && sv->index >= 0 && sweep[sv->index].acc >= 0) {
        const wva_grid_best& g = sweep[sv->index];
        out->HaveSweep = true;
        out->SweepAccelerator = system.AcceleratorNames()[(size_t)g.acc];
        out->SweepReplicas = g.replicas; out->SweepBatch = g.batch; out->SweepCost = g.cost; out->SweepITL = g.itl; out->SweepTTFT = g.ttft; out->SweepRho = g.rho;
    }
    return true;
}
inline std::string ToJSON(const OptimizedAllocExt& a) {
    std::ostringstream o;
    o << "{\"accelerator\":" << JsonString(a.Accelerator) << ",\"numReplicas\":" << a.NumReplicas
      << ",\"maxBatch\":" << a.MaxBatch << ",\"cost\":" << JsonFloat32(a.Cost) << ",\"itlAverage\":" << JsonFloat32(a.ITLAverage)
      << ",\"ttftAverage\":" << JsonFloat32(a.TTFTAverage) << ",\"rho\":" << JsonFloat32(a.Rho)
      << ",\"maxArrvRatePerReplica\":" << JsonFloat32(a.MaxArrvRatePerReplica);
    if (a.HaveSweep)
        o << ",\"sweep\":{\"accelerator\":" << JsonString(a.SweepAccelerator) << ",\"numReplicas\":" << a.SweepReplicas << ",\"maxBatch\":" << a.SweepBatch
          << ",\"cost\":" << JsonFloat32(a.SweepCost) << ",\"itlAverage\":" << JsonFloat32(a.SweepITL) << ",\"ttftAverage\":" << JsonFloat32(a.SweepTTFT)
          << ",\"rho\":" << JsonFloat32(a.SweepRho) << "}";
    o << "}";
    return o.str();
}

}  // namespace utils
}  // namespace wva
