// spec.h -- front end, tier (a): recognise an energy specification (.t file).
//
// The reference executes the .t file as Lua/Terra and differentiates it symbolically
// (API/src/o.t:840-882, :2425-2460).  This tier does not interpret the DSL: it tokenises the file
// (comments and whitespace removed), hashes the canonical token stream and looks the hash up in a
// registry of the energies that have hand-written kernel sets.  A spec whose tokens differ from every
// registered one is REJECTED (Opt_ProblemPlan returns NULL) rather than silently solved as
// something else.  The general front end (Lua-subset interpreter + AD + NVRTC) is SURVEY.md 8(f) rank 1.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace optb {

struct SpecMatch {
    bool ok = false;
    std::string energy;       // registry name, e.g. "image_warping"
    std::string message;      // why not ok
    uint64_t tokenHash = 0;
    size_t tokenCount = 0;
};

// canonical token stream of Lua source (comments/whitespace stripped, number literals normalised)
std::vector<std::string> tokenizeLua(const std::string& src);
uint64_t hashTokens(const std::vector<std::string>& toks);
SpecMatch matchSpecFile(const std::string& filename);

}  // namespace optb
