// Statically compiled elementwise programs: one straight-line kernel per catalogued program and rank class.
// Compiled EW_NSHARDS times (-DEW_SHARD=k) so the catalogue builds in parallel; shard 0 also owns the registry.
#include <vector>

#include "ewise_kernel.cuh"

#ifndef EW_SHARD
#define EW_SHARD -1   // single translation unit: everything
#endif

namespace adb {

std::vector<EwCatalogEntry>& ew_registry();

namespace {
struct EwRegistrar {
    EwRegistrar(int n, const int* sig, EwLaunchFn f1, EwLaunchFn f2, EwLaunchFn fn) {
        EwCatalogEntry e;
        e.n_instr = n; e.sig = sig; e.fn[0] = f1; e.fn[1] = f2; e.fn[2] = fn;
        ew_registry().push_back(e);
    }
};
}  // namespace

#define EW_ENTRY(ID, N)                                                                                              \
    static EwRegistrar ew_reg##ID(N, ew_sig##ID, ew_launch<1, EwStatic<EwP##ID>>, ew_launch<2, EwStatic<EwP##ID>>, \
                                  ew_launch<EW_RANK, EwStatic<EwP##ID>>);
#include "ewise_catalog.inc"
#undef EW_ENTRY

#if EW_SHARD <= 0
std::vector<EwCatalogEntry>& ew_registry() {
    static std::vector<EwCatalogEntry> r;
    return r;
}
const EwCatalogEntry* ew_catalog(int* count) {
    std::vector<EwCatalogEntry>& r = ew_registry();
    *count = (int)r.size();
    return r.data();
}
#endif

}  // namespace adb
