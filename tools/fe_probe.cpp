// fe_probe -- runs the .t front end on files and prints what it found (debugging aid / CPU test helper)
#include <cstdio>
#include <cstring>
#include "../opt_b200/csrc/frontend.h"
int main(int argc, char** argv) {
    bool dump = false; int rc = 0; const char* emit = nullptr;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--dump")) { dump = true; continue; }
        if (!strcmp(argv[i], "--emit") && i + 1 < argc) { emit = argv[++i]; continue; }   // write the generated source of the next spec to this file
        auto cs = optb::fe::compileSpecFile(argv[i]);
        if (!cs.ok) { printf("FAIL %s: %s\n", argv[i], cs.error.c_str()); rc = 1; continue; }
        int nu = 0; for (auto& im : cs.images) nu += im.unknown;
        printf("OK   %s: %zu images (%d unknown), %zu params, %zu graphs, %zu scalar residual groups, pre=%d, %zu bytes of CUDA\n", argv[i], cs.images.size(), nu,
               cs.params.size(), cs.graphs.size(), cs.groups.size(), (int)cs.usePreconditioner, cs.cudaSource.size());
        if (dump) printf("%s\n", cs.cudaSource.c_str());
        if (emit) { FILE* f = fopen(emit, "w"); if (!f) { printf("FAIL cannot write %s\n", emit); rc = 1; } else { fputs(cs.cudaSource.c_str(), f); fclose(f); } emit = nullptr; }
    }
    return rc;
}
