-- robust non-rigid surface alignment (opt_b200 test spec; same energy and parameter positions as the
-- robust_nonrigid_alignment example of niessner/Opt).
--   params: 0 w_fitSqrt  1 w_regSqrt  2 Offset  3 Angle  4 RobustWeights (unknowns)
--           5 UrShape  6 Constraints  7 ConstraintNormals  8 edge count  9 v0  10 v1
local N = Dim("N", 0)
local w_fitSqrt = Param("w_fitSqrt", float, 0)
local w_regSqrt = Param("w_regSqrt", float, 1)
local w_confSqrt = 0.1
local Offset            = Unknown("Offset", opt_float3, { N }, 2)
local Angle             = Unknown("Angle", opt_float3, { N }, 3)
local RobustWeights     = Unknown("RobustWeights", opt_float, { N }, 4)
local UrShape           = Array("UrShape", opt_float3, { N }, 5)
local Constraints       = Array("Constraints", opt_float3, { N }, 6)
local ConstraintNormals = Array("ConstraintNormals", opt_float3, { N }, 7)
local G = Graph("G", 8, "v0", { N }, 9, "v1", { N }, 10)
UsePreconditioner(true)

local weight = RobustWeights(0)
local hasTarget = greatereq(Constraints(0), -999999.9)

-- point-to-plane distance to the correspondence, scaled by the robust weight
local planeDist = weight * ConstraintNormals(0):dot(Offset(0) - Constraints(0))
Energy(w_fitSqrt * Select(hasTarget, planeDist, 0.0))

-- keep the robust weights near one
Energy(w_confSqrt * Select(hasTarget, 1 - (weight * weight), 0.0))

-- as-rigid-as-possible coupling along the mesh edges
local restEdge = UrShape(G.v0) - UrShape(G.v1)
Energy(w_regSqrt * ((Offset(G.v0) - Offset(G.v1)) - Rotate3D(Angle(G.v0), restEdge)))
