-- embedded deformation with per-node affine matrices (opt_b200 test spec; same energy and parameter positions as
-- the embedded_mesh_deformation example of niessner/Opt).  Served by the general front end.
--   params: 0 w_fitSqrt  1 w_regSqrt  2 w_rotSqrt  3 Offset  4 RotMatrix (unknowns)  5 UrShape  6 Constraints  7 edge count  8 v0  9 v1
local N = Dim("N", 0)
local w_fitSqrt   = Param("w_fitSqrt", float, 0)
local w_regSqrt   = Param("w_regSqrt", float, 1)
local w_rotSqrt   = Param("w_rotSqrt", float, 2)
local Offset      = Unknown("Offset", opt_float3, { N }, 3)
local RotMatrix   = Unknown("RotMatrix", opt_float9, { N }, 4)
local UrShape     = Image("UrShape", opt_float3, { N }, 5)
local Constraints = Image("Constraints", opt_float3, { N }, 6)
local G = Graph("G", 7, "v0", { N }, 8, "v1", { N }, 9)
UsePreconditioner(true)

local pinned = greatereq(Constraints(0)(0), -999999.9)
Energy(Select(pinned, w_fitSqrt * (Offset(0) - Constraints(0)), 0))

-- the 3x3 matrix of every node should stay a rotation: orthonormal columns
local M = RotMatrix(0)
local col = { Vector(M(0), M(3), M(6)), Vector(M(1), M(4), M(7)), Vector(M(2), M(5), M(8)) }
Energy(w_rotSqrt * Dot3(col[1], col[2]))
Energy(w_rotSqrt * Dot3(col[1], col[3]))
Energy(w_rotSqrt * Dot3(col[2], col[3]))
for k = 1, 3 do
    Energy(w_rotSqrt * (Dot3(col[k], col[k]) - 1))
end

local moved = (Offset(G.v1) - Offset(G.v0)) - Matrix3x3Mul(RotMatrix(G.v0), (UrShape(G.v1) - UrShape(G.v0)))
Energy(w_regSqrt * moved)
