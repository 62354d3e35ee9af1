-- Mesh fairing with cotangent weights (test specification of opt_b200; the parameter table matches the
-- cotangent_mesh_smoothing example of niessner/Opt).  One hyperedge per mesh edge (a, b) with the two vertices c, d
-- opposite to it in the adjacent triangles.  No built-in kernel set: general front end.
--   slot 0 / 1: square roots of the data and fairness weights     slot 2: vertex positions, UNKNOWN     slot 3: input positions
--   slot 4: number of hyperedges      slots 5..8: vertex indices a, b, c, d
local verts = { Dim("N", 0) }
local sqrtData, sqrtFair = Param("w_fit", float, 0), Param("w_reg", float, 1)
local pos   = Unknown("X", opt_float3, verts, 2)
local input = Array("A", opt_float3, verts, 3)
local wing  = Graph("G", 4, "v0", verts, 5, "v1", verts, 6, "v2", verts, 7, "v3", verts, 8)
UsePreconditioner(true)

-- stay near the input surface
Energy(sqrtData * (pos(0) - input(0)))

-- cot of the angle between two unit vectors, guarded against degenerate (zero-area) corners
local tiny = 0.0001
local function cotBetween(s, t)
    local c = Dot3(s, t)
    local sin2 = Dot3(s, s) * Dot3(t, t) - c * c
    return c / Sqrt(Select(greater(sin2, 0.0), sin2, tiny))
end
local function unitFrom(apex, tip) return normalize(pos(tip) - pos(apex)) end

local a, b, c, d = wing.v0, wing.v1, wing.v2, wing.v3
local halfSum = 0.5 * (cotBetween(unitFrom(c, a), unitFrom(c, b)) + cotBetween(unitFrom(d, a), unitFrom(d, b)))
local edgeWeight = Sqrt(Select(greater(halfSum, 0.0), halfSum, tiny))
Energy(sqrtFair * edgeWeight * (pos(b) - pos(a)))
