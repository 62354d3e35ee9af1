// fe_probe -- runs the .t front end on files and prints what it found (debugging aid / CPU test helper)
#include <cstdio>
#include <cstring>
#include "../opt_b200/csrc/frontend.h"
int main(int argc, char** argv) {
    bool dump = false, meta = false; int rc = 0; const char* emit = nullptr;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--dump")) { dump = true; continue; }
        if (!strcmp(argv[i], "--meta")) { meta = true; continue; }   // machine-readable description of the parameter table
        if (!strcmp(argv[i], "--emit") && i + 1 < argc) { emit = argv[++i]; continue; }   // write the generated source of the next spec to this file
        auto cs = optb::fe::compileSpecFile(argv[i]);
        if (!cs.ok) { printf("FAIL %s: %s\n", argv[i], cs.error.c_str()); rc = 1; continue; }
        int nu = 0; for (auto& im : cs.images) nu += im.unknown;
        printf("OK   %s: %zu images (%d unknown), %zu params, %zu graphs, %zu scalar residual groups, pre=%d, %zu bytes of CUDA\n", argv[i], cs.images.size(), nu,
               cs.params.size(), cs.graphs.size(), cs.groups.size(), (int)cs.usePreconditioner, cs.cudaSource.size());
        if (dump) printf("%s\n", cs.cudaSource.c_str());
        if (meta) {
            static const char* tn[] = { "optfloat", "float", "double", "uint8", "int32" };
            printf("MAXDIM %d\n", cs.maxDimIndex);
            for (auto& im : cs.images) { printf("IMAGE %s %s %d %d %d", im.name.c_str(), tn[(int)im.type], im.channels, im.paramIndex, (int)im.unknown); for (int d : im.dims) printf(" %d", d); printf("\n"); }
            for (auto& pm : cs.params) printf("PARAM %s %s %d\n", pm.name.c_str(), tn[(int)pm.type], pm.paramIndex);
            for (auto& g : cs.graphs) { printf("GRAPH %s %d %zu\n", g.name.c_str(), g.countIndex, g.slots.size()); for (auto& sl : g.slots) { printf("SLOT %s %d", sl.name.c_str(), sl.paramIndex); for (int d : sl.dims) printf(" %d", d); printf("\n"); } }
        }
        if (emit) { FILE* f = fopen(emit, "w"); if (!f) { printf("FAIL cannot write %s\n", emit); rc = 1; } else { fputs(cs.cudaSource.c_str(), f); fclose(f); } emit = nullptr; }
    }
    return rc;
}
