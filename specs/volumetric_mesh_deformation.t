-- ARAP deformation of a regular 3-D lattice (opt_b200 test spec; same energy and parameter positions as the
-- volumetric_mesh_deformation example of niessner/Opt).  Not in the kernel registry: served by the general front end.
--   params: 0 Offset  1 Angle (unknowns)  2 UrShape  3 Constraints  4 w_fitSqrt  5 w_regSqrt
local W = Dim("W", 0)
local H = Dim("H", 1)
local D = Dim("D", 2)
local lattice = { W, H, D }
local Offset      = Unknown("Offset", opt_float3, lattice, 0)
local Angle       = Unknown("Angle", opt_float3, lattice, 1)
local UrShape     = Array("UrShape", opt_float3, lattice, 2)
local Constraints = Array("Constraints", opt_float3, lattice, 3)
local w_fitSqrt   = Param("w_fitSqrt", float, 4)
local w_regSqrt   = Param("w_regSqrt", float, 5)
UsePreconditioner(true)

local pinned = greatereq(Constraints(0, 0, 0)(0), -999999.9)
Energy(Select(pinned, w_fitSqrt * (Offset(0, 0, 0) - Constraints(0, 0, 0)), 0))

local here, rest, turn = Offset(0, 0, 0), UrShape(0, 0, 0), Angle(0, 0, 0)
for a, b, c in Stencil { { 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 } } do
    local rigid = (here - Offset(a, b, c)) - Rotate3D(turn, rest - UrShape(a, b, c))
    Energy(w_regSqrt * Select(InBounds(0, 0, 0), Select(InBounds(a, b, c), rigid, 0.0), 0.0))
end
