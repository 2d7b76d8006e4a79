// host_test.cpp — runs the reference's reconcile call sequence (variantautoscaling_controller.go:143-166)
// through the C++ host mirror and prints one JSON object per scenario for tests/test_host_cpp.py.
// Scenarios restate the reference's own integration tests (internal/optimizer/optimizer_test.go:245-457).
#include <cstdio>
#include <cstring>
#include "wva_host.hpp"
#include "wva_adapters.hpp"

using namespace wva;

static config::SystemSpec optimizerFixture(float arrivalRpm, int outTokens, float sloItl, float sloTtft, const char* acc, float cost) {
    config::SystemSpec sp;
    config::AcceleratorSpec a; a.Name = acc; a.Type = acc; a.Multiplicity = 1; a.Cost = cost; sp.Accelerators.push_back(a);
    config::ModelAcceleratorPerfData pd; pd.Name = "m"; pd.Acc = acc; pd.AccCount = 1; pd.MaxBatchSize = 4; pd.AtTokens = 0;
    pd.Decode.Alpha = 20.28f; pd.Decode.Beta = 0.72f; sp.Models.push_back(pd);
    config::ServiceClassSpec sc; sc.Name = "Premium"; sc.Priority = 1; sc.ModelTargets.push_back({"m", sloItl, sloTtft, 0.0f}); sp.ServiceClasses.push_back(sc);
    config::ServerSpec sv; sv.Name = "va:default"; sv.Class = "Premium"; sv.Model = "m"; sv.KeepAccelerator = true; sv.MinNumReplicas = 1; sv.MaxBatchSize = 4;
    sv.CurrentAlloc.Accelerator = acc; sv.CurrentAlloc.NumReplicas = 1; sv.CurrentAlloc.Cost = cost;
    sv.CurrentAlloc.Load.ArrivalRate = arrivalRpm; sv.CurrentAlloc.Load.AvgInTokens = 20; sv.CurrentAlloc.Load.AvgOutTokens = outTokens;
    sp.Servers.push_back(sv);
    sp.Optimizer.Unlimited = true;            // utils.CreateSystemData hard-codes it (internal/utils/utils.go:170-173)
    return sp;
}

static void reconcile(NativeContext& native, const config::SystemSpec& spec, const char* label) {
    core::System system(native);                                             // inferno.NewSystem()
    const config::OptimizerSpec& os = system.SetFromSpec(spec);              // system.SetFromSpec(&systemData.Spec)
    solver::Optimizer optimizer(os);                                         // infernoSolver.NewOptimizerFromSpec
    manager::Manager manager(system, optimizer);                             // infernoManager.NewManager
    modelanalyzer::ModelAnalyzer analyzer(system);                           // analyzer.NewModelAnalyzer(system)
    size_t nAlloc = 0;
    std::vector<optimizer::VariantRef> vas;
    for (const auto& kv : system.Servers()) {
        const std::string& full = kv.first; size_t c = full.find(':');
        std::string name = full.substr(0, c), ns = c == std::string::npos ? "" : full.substr(c + 1);
        auto resp = analyzer.AnalyzeModel(name, ns);                         // modelAnalyzer.AnalyzeModel(ctx, va)
        nAlloc += resp.Allocations.size();
        vas.push_back({name, ns});
    }
    optimizer::VariantAutoscalingsEngine engine(manager, system);           // NewVariantAutoscalingsEngine
    std::printf("{\"scenario\": \"%s\", \"analyze_allocations\": %zu, ", label, nAlloc);
    try {
        auto out = engine.Optimize(vas);                                     // engine.Optimize(ctx, vaList, responses)
        std::printf("\"error\": null, \"optimized\": {");
        bool first = true;
        for (const auto& kv : out) {
            auto sv = system.GetServer(kv.first + ":default");
            std::printf("%s\"%s\": {\"accelerator\": \"%s\", \"replicas\": %lld, \"itl\": %.9g, \"ttft\": %.9g, \"cost\": %.9g}", first ? "" : ", ",
                        kv.first.c_str(), kv.second.Accelerator.c_str(), (long long)kv.second.NumReplicas,
                        sv && sv->Allocation() ? sv->Allocation()->ITL() : 0.0f, sv && sv->Allocation() ? sv->Allocation()->TTFT() : 0.0f,
                        sv && sv->Allocation() ? sv->Allocation()->Cost() : 0.0f);
            first = false;
        }
        std::printf("}, \"by_type\": {");
        first = true;
        for (const auto& kv : system.AllocationByTypeMap()) { std::printf("%s\"%s\": {\"count\": %lld, \"cost\": %.9g}", first ? "" : ", ", kv.first.c_str(), (long long)kv.second.count, kv.second.cost); first = false; }
        std::printf("}}\n");
    } catch (const Error& e) {
        std::printf("\"error\": {\"code\": %d, \"message\": \"%s\"}}\n", e.code, e.what());
    }
}

int main() {
    try {
        NativeContext native(0);
        reconcile(native, optimizerFixture(0.0f, 0, 80.0f, 500.0f, "A100", 40.0f), "no_load");            // optimizer_test.go:245-335
        reconcile(native, optimizerFixture(1200.0f, 200, 80.0f, 500.0f, "A100", 40.0f), "scale_out");     // :337-457
        reconcile(native, optimizerFixture(1200.0f, 200, 1.0f, 1.0f, "A100", 40.0f), "infeasible_slo");   // Optimize returns the error of :38-40
        // two accelerators, limited capacity, greedy + PriorityExhaustive
        config::SystemSpec sp = optimizerFixture(1200.0f, 200, 80.0f, 500.0f, "A100", 40.0f);
        config::AcceleratorSpec h; h.Name = "H100"; h.Type = "H100"; h.Multiplicity = 1; h.Cost = 100.0f; sp.Accelerators.push_back(h);
        config::ModelAcceleratorPerfData pd = sp.Models[0]; pd.Acc = "H100"; pd.Decode.Alpha = 7.47f; pd.Decode.Beta = 0.044f; pd.MaxBatchSize = 4; sp.Models.push_back(pd);
        sp.Servers[0].KeepAccelerator = false;
        sp.Capacity.push_back({"A100", 10}); sp.Capacity.push_back({"H100", 8});
        sp.Optimizer.Unlimited = false; sp.Optimizer.SaturationPolicy = "PriorityExhaustive";
        reconcile(native, sp, "limited_greedy");
        // Allocation.Scale / ReAllocate (allocation.go:165-207) on the two-accelerator system with keepAccelerator:
        // Server.Calculate only sizes the current accelerator, ReAllocate looks at all of them
        {
            config::SystemSpec sk = sp;
            sk.Servers[0].KeepAccelerator = true;
            sk.Optimizer.Unlimited = true;
            core::System system(native);
            system.SetFromSpec(sk);
            system.Calculate();
            const auto sv = system.GetServer("va:default");
            core::Allocation cur; cur.accelerator = "A100"; cur.numReplicas = 1;
            auto sc = system.Scale(cur, "va:default");
            auto ra = system.ReAllocate("va:default");
            auto gone = system.Scale(cur, "nobody");
            core::Allocation odd; odd.accelerator = "L40S";
            auto noacc = system.Scale(odd, "va:default");
            std::printf("{\"scenario\": \"scale_realloc\", \"candidates\": %zu, \"scale\": {\"replicas\": %lld, \"inc\": %lld, \"cost\": %.9g, \"value\": %.9g}, "
                        "\"realloc\": {\"accelerator\": \"%s\", \"replicas\": %lld, \"cost\": %.9g}, \"nil_cases\": %d, \"candidates_after\": %zu}\n",
                        sv->AllAllocations().size(), sc.first ? (long long)sc.first->NumReplicas() : -1LL, (long long)sc.second,
                        sc.first ? sc.first->Cost() : 0.0f, sc.first ? sc.first->Value() : 0.0f,
                        ra.second.c_str(), ra.first ? (long long)ra.first->NumReplicas() : -1LL, ra.first ? ra.first->Cost() : 0.0f,
                        (int)(!gone.first && gone.second == 0) + (int)(!noacc.first && noacc.second == 0),
                        (system.Calculate(), sv->AllAllocations().size()));
        }
        // incremental updates (system.go:99-171): add a second server, re-optimize, remove it again
        {
            core::System system(native);
            const config::OptimizerSpec& os = system.SetFromSpec(optimizerFixture(1200.0f, 200, 80.0f, 500.0f, "A100", 40.0f));
            solver::Optimizer optimizer(os);
            manager::Manager manager(system, optimizer);
            config::ServerSpec vb = system.GetServer("va:default")->spec;
            vb.Name = "vb:default"; vb.CurrentAlloc.Load.ArrivalRate = 600.0f;
            system.AddServerFromSpec(vb);
            system.Calculate(); manager.Optimize();
            const long long n2 = (long long)system.Servers().size();
            const long long ra2 = system.GetServer("va:default")->Allocation() ? (long long)system.GetServer("va:default")->Allocation()->NumReplicas() : -1;
            const long long rb2 = system.GetServer("vb:default")->Allocation() ? (long long)system.GetServer("vb:default")->Allocation()->NumReplicas() : -1;
            const bool removed = system.RemoveServer("vb:default"), again = system.RemoveServer("vb:default");
            system.Calculate(); manager.Optimize();
            const long long n1 = (long long)system.Servers().size();
            const long long ra1 = system.GetServer("va:default")->Allocation() ? (long long)system.GetServer("va:default")->Allocation()->NumReplicas() : -1;
            std::printf("{\"scenario\": \"incremental\", \"servers_after_add\": %lld, \"va_after_add\": %lld, \"vb_after_add\": %lld, "
                        "\"removed\": %d, \"removed_again\": %d, \"servers_after_remove\": %lld, \"va_after_remove\": %lld}\n",
                        n2, ra2, rb2, (int)removed, (int)again, n1, ra1);
        }
        // delta uploads (system.go:99-171 on the resident image) and Solve -> ReAllocate -> AllocateByType on one System
        {
            config::SystemSpec big = optimizerFixture(1200.0f, 200, 80.0f, 500.0f, "A100", 40.0f);
            config::AcceleratorSpec h2; h2.Name = "H100"; h2.Type = "H100"; h2.Multiplicity = 1; h2.Cost = 100.0f; big.Accelerators.push_back(h2);
            config::ModelAcceleratorPerfData pd2 = big.Models[0]; pd2.Acc = "H100"; pd2.Decode.Alpha = 7.47f; pd2.Decode.Beta = 0.044f; big.Models.push_back(pd2);
            for (int i = 0; i < 400; ++i) { config::ServerSpec sv = big.Servers[0]; sv.Name = "s" + std::to_string(i) + ":default"; sv.CurrentAlloc.Load.ArrivalRate = 60.0f + 3.0f * (float)i; big.Servers.push_back(sv); }
            core::System system(native);
            const config::OptimizerSpec& os = system.SetFromSpec(big);
            const long long fullBytes = (long long)system.UploadBytes();
            solver::Optimizer optimizer(os);
            manager::Manager manager(system, optimizer);
            system.Calculate(); manager.Optimize();
            const long long before = (long long)system.GetServer("s7:default")->Allocation()->NumReplicas();
            config::ServerSpec s7 = system.GetServer("s7:default")->spec; s7.CurrentAlloc.Load.ArrivalRate = 2400.0f;
            system.AddServerFromSpec(s7);                                   // replaces the existing server: ONE row goes up
            const long long deltaBytes = (long long)system.UploadBytes();
            system.SetCountFromSpec({"A100", 100000});
            const long long capBytes = (long long)system.UploadBytes();
            system.Calculate(); manager.Optimize();
            const long long after = (long long)system.GetServer("s7:default")->Allocation()->NumReplicas();
            // ReAllocate / Scale after Solve must leave the solved state alone: AllocateByType still works
            auto ra = system.ReAllocate("s7:default");
            system.AllocateByType();
            long long total = 0; for (const auto& kv : system.AllocationByTypeMap()) total += (long long)kv.second.count;
            std::printf("{\"scenario\": \"delta\", \"full_bytes\": %lld, \"delta_bytes\": %lld, \"capacity_bytes\": %lld, \"replicas_before\": %lld, "
                        "\"replicas_after\": %lld, \"realloc\": \"%s\", \"by_type_total\": %lld, \"servers\": %zu}\n",
                        fullBytes, deltaBytes, capBytes, before, after, ra.second.c_str(), total, system.Servers().size());
        }
        // the formats either side of the path: ConfigMap-shaped strings in (internal/utils/utils.go:108-311), JSON out.
        // Inputs: the reference's own chart values for Llama-3.1-8B on L40S (BASELINE config 1).
        {
            std::map<std::string, std::map<std::string, std::string>> accCm = {{"L40S", {{"device", "NVIDIA-L40S"}, {"cost", "32.00"}}},
                                                                                  {"broken", {{"device", "x"}, {"cost", "n/a"}}}};
            utils::ServiceClass premium; premium.Name = "Premium"; premium.Priority = 1; premium.Data.push_back({"default/default", 24.0, 500.0});
            config::SystemSpec sd = utils::CreateSystemData(accCm, {premium});
            utils::AcceleratorProfile prof; prof.Acc = "L40S"; prof.AccCount = 1; prof.MaxBatchSize = 512;
            prof.DecodeParms = {{"alpha", "22.619"}, {"beta", "0.181"}}; prof.PrefillParms = {{"gamma", "226.19"}, {"delta", "0.018"}};
            const std::string e1 = utils::AddModelAcceleratorProfileToSystemData(sd, "default/default", prof);
            utils::AcceleratorProfile bad = prof; bad.DecodeParms = {{"alpha", "abc"}, {"beta", "1"}};
            const std::string e2 = utils::AddModelAcceleratorProfileToSystemData(sd, "default/default", bad);
            utils::VariantAutoscaling va; va.Name = "llama"; va.Namespace = "default"; va.ModelID = "default/default"; va.AcceleratorLabel = "L40S";
            va.Accelerators.push_back(prof);
            va.CurrentAlloc.Accelerator = "L40S"; va.CurrentAlloc.NumReplicas = 1; va.CurrentAlloc.MaxBatch = 512; va.CurrentAlloc.VariantCost = "32.00";
            va.CurrentAlloc.ITLAverage = "NaN"; va.CurrentAlloc.TTFTAverage = ""; va.CurrentAlloc.Load.ArrivalRate = "600"; va.CurrentAlloc.Load.AvgInputTokens = "128.7";
            va.CurrentAlloc.Load.AvgOutputTokens = "128";
            utils::AddServerInfoToSystemData(sd, va, "Premium", false);
            core::System system(native);
            const config::OptimizerSpec& os = system.SetFromSpec(sd);
            solver::Optimizer optimizer(os);
            manager::Manager manager(system, optimizer);
            system.Calculate(); manager.Optimize();
            std::vector<wva_grid_best> sweep(system.Servers().size());
            native.check(wva_analyze_grid(native.get(), 8, 256, sweep.data(), nullptr, nullptr));
            utils::OptimizedAllocExt ext;
            const bool ok = utils::CreateOptimizedAllocExt(system, "llama", "default", sweep.data(), &ext);
            utils::OptimizedAllocExt none;
            const bool missing = utils::CreateOptimizedAllocExt(system, "nobody", "default", sweep.data(), &none);
            std::printf("{\"scenario\": \"adapters\", \"accelerators\": %zu, \"err_ok\": \"%s\", \"err_bad\": \"%s\", \"in_tokens\": %d, \"itl_in\": %.9g, "
                        "\"min_replicas\": %d, \"server_batch\": %d, \"found\": %d, \"missing\": %d, \"solution\": %s, \"optimized\": %s, "
                        "\"floats\": [\"%s\", \"%s\", \"%s\", \"%s\", \"%s\", \"%s\"]}\n",
                        sd.Accelerators.size(), e1.c_str(), e2.c_str(), sd.Servers[0].CurrentAlloc.Load.AvgInTokens, sd.Servers[0].CurrentAlloc.ITLAverage,
                        sd.Servers[0].MinNumReplicas, sd.Servers[0].MaxBatchSize, (int)ok, (int)missing,
                        utils::ToJSON(system.GenerateSolution()).c_str(), utils::ToJSON(ext).c_str(),
                        utils::JsonFloat32(1e-7f).c_str(), utils::JsonFloat32(1e21f).c_str(), utils::JsonFloat32(0.1f).c_str(),
                        utils::JsonFloat32(16777216.0f).c_str(), utils::JsonFloat32(-2.5e-5f).c_str(), utils::JsonFloat32(3.4028235e38f).c_str());
        }
    } catch (const Error& e) {
        std::printf("{\"fatal\": {\"code\": %d, \"message\": \"%s\"}}\n", e.code, e.what());
        return 1;
    }
    return 0;
}
