// host_test.cpp — runs the reference's reconcile call sequence (variantautoscaling_controller.go:143-166)
// through the C++ host mirror and prints one JSON object per scenario for tests/test_host_cpp.py.
// Scenarios restate the reference's own integration tests (internal/optimizer/optimizer_test.go:245-457).
#include <cstdio>
#include <cstring>
#include "wva_host.hpp"

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
    } catch (const Error& e) {
        std::printf("{\"fatal\": {\"code\": %d, \"message\": \"%s\"}}\n", e.code, e.what());
        return 1;
    }
    return 0;
}
