-- 3-D as-rigid-as-possible mesh deformation over a vertex graph (opt_b200 test spec; same energy
-- and parameter positions as the arap_mesh_deformation example of niessner/Opt).
--   params: 0 w_fitSqrt  1 w_regSqrt  2 Offset (unknown)  3 Angle (unknown)  4 UrShape  5 Constraints
--           6 edge count  7 v0 indices  8 v1 indices
local N = Dim("N", 0)
local w_fitSqrt   = Param("w_fitSqrt", float, 0)
local w_regSqrt   = Param("w_regSqrt", float, 1)
local Offset      = Unknown("Offset", opt_float3, { N }, 2)
local Angle       = Unknown("Angle", opt_float3, { N }, 3)
local UrShape     = Array("UrShape", opt_float3, { N }, 4)
local Constraints = Array("Constraints", opt_float3, { N }, 5)
local G = Graph("G", 6, "v0", { N }, 7, "v1", { N }, 8)
UsePreconditioner(true)

-- handle vertices: first channel below -999999.9 means "free vertex"
local isHandle = greatereq(Constraints(0, 0), -999999.9)
Energy(Select(isHandle, w_fitSqrt * (Offset(0) - Constraints(0)), 0))

-- per directed edge: rotated rest edge must match the deformed edge
local restEdge = UrShape(G.v0) - UrShape(G.v1)
local nowEdge = Offset(G.v0) - Offset(G.v1)
Energy(w_regSqrt * (nowEdge - Rotate3D(Angle(G.v0), restEdge)))
