-- cotangent-Laplacian mesh smoothing over 4-vertex hyperedges (opt_b200 test spec; same energy and parameter
-- positions as the cotangent_mesh_smoothing example of niessner/Opt).  Served by the general front end.
--   params: 0 w_fit  1 w_reg  2 X (unknown)  3 A  4 edge count  5 v0  6 v1  7 v2  8 v3
local N = Dim("N", 0)
local w_fitSqrt = Param("w_fit", float, 0)
local w_regSqrt = Param("w_reg", float, 1)
local X = Unknown("X", opt_float3, { N }, 2)
local A = Array("A", opt_float3, { N }, 3)
local G = Graph("G", 4, "v0", { N }, 5, "v1", { N }, 6, "v2", { N }, 7, "v3", { N }, 8)
UsePreconditioner(true)

local function cotangent(p, q)
    local pq = Dot3(p, q)
    local area2 = Dot3(p, p) * Dot3(q, q) - pq * pq
    area2 = Select(greater(area2, 0.0), area2, 0.0001)
    return pq / Sqrt(area2)
end

Energy(w_fitSqrt * (X(0) - A(0)))

local e02 = normalize(X(G.v0) - X(G.v2))
local e12 = normalize(X(G.v1) - X(G.v2))
local e03 = normalize(X(G.v0) - X(G.v3))
local e13 = normalize(X(G.v1) - X(G.v3))
local weight = 0.5 * (cotangent(e02, e12) + cotangent(e03, e13))
weight = Sqrt(Select(greater(weight, 0.0), weight, 0.0001))
Energy(w_regSqrt * weight * (X(G.v1) - X(G.v0)))
