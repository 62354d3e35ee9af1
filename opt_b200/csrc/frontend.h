// frontend.h -- general energy-specification front end.
//
// The reference runs a .t file as Lua/Terra code inside an environment of DSL functions (API/src/lib.t:1-133,
// API/src/o.t:930-1115), differentiates the resulting expressions symbolically (ad.t) and JIT-compiles Terra
// kernels.  Here a small Lua-subset interpreter (C++) runs the same file against the same DSL surface and
// records every Energy() term as an expression DAG; code generation then prints ONE device function per
// problem -- `eval`, the residuals exactly as written -- which NVRTC compiles together with
// generic_device.cuh.  Differentiation is not symbolic: the generic engine evaluates `eval` with forward-mode
// dual numbers on the device (generic_engine.cuh).  Specs whose token hash is in the registry still take
// their hand-written / statically compiled kernel sets; this path serves everything else
// (or everything, with OPTB_FRONTEND=1).
#pragma once
#include <string>
#include <vector>

namespace optb {
namespace fe {

enum class ScalarType { OptFloat, Float, Double, UInt8, Int32 };

struct ImageDecl {          // Unknown(...) / Array(...) / Image(...)
    std::string name;
    ScalarType type = ScalarType::OptFloat;
    int channels = 1;
    std::vector<int> dims;  // indices into the `dimensions` argument of Opt_ProblemPlan, fastest first
    int paramIndex = -1;    // position in problemparams
    bool unknown = false;
    int unknownId = -1;     // index among the unknown arrays (spec order)
};
struct ParamDecl { std::string name; ScalarType type = ScalarType::Float; int paramIndex = -1; };
struct GraphDecl {
    std::string name; int countIndex = -1;
    struct Slot { std::string name; std::vector<int> dims; int paramIndex = -1; };
    std::vector<Slot> slots;
};
struct GroupDecl {          // one scalar residual over its domain
    bool centered = true;
    std::vector<int> dims;  // centred: the index space
    int graph = -1;         // graph: which graph
    int centerUnknown = 0;  // unknown array whose index space is the centre (Exclude test)
    int ndeps = 1;
};

struct CompiledSpec {
    bool ok = false;
    std::string error;
    std::vector<ImageDecl> images;
    std::vector<ParamDecl> params;
    std::vector<GraphDecl> graphs;
    std::vector<GroupDecl> groups;
    bool usePreconditioner = false;   // default false (o.t:214)
    int maxDimIndex = -1;
    std::string cudaSource;           // definition of  template<class T> struct GenSpec { eval; excluded; }
};

// Runs `filename` (Lua subset + Opt DSL) and generates the device code.  Pure host code: no CUDA needed.
CompiledSpec compileSpecFile(const std::string& filename);
CompiledSpec compileSpecSource(const std::string& source, const std::string& chunkName);

}  // namespace fe
}  // namespace optb
