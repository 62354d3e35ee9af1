// spec_hash -- prints the canonical token hash of .t files (see opt_b200/csrc/spec.h).
#include <cstdio>
#include <fstream>
#include <sstream>
#include "../opt_b200/csrc/spec.h"
int main(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
        std::ifstream f(argv[i], std::ios::binary);
        if (!f) { fprintf(stderr, "cannot open %s\n", argv[i]); return 1; }
        std::stringstream ss; ss << f.rdbuf();
        auto t = optb::tokenizeLua(ss.str());
        printf("%016llx %zu %s\n", (unsigned long long)optb::hashTokens(t), t.size(), argv[i]);
    }
    return 0;
}
