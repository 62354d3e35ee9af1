-- As-rigid-as-possible deformation of a regular 3-D lattice of control points (test specification of opt_b200; the
-- parameter table matches the volumetric_mesh_deformation example of niessner/Opt).  No built-in kernel set: general front end.
--   slot 0: lattice point positions, UNKNOWN      slot 1: lattice point Euler angles, UNKNOWN      slot 2: rest positions
--   slot 3: handle targets (x below -999999.9 = free point)      slot 4 / 5: square roots of the handle and rigidity weights
local cells = { Dim("W", 0), Dim("H", 1), Dim("D", 2) }
local pos    = Unknown("Offset", opt_float3, cells, 0)
local euler  = Unknown("Angle", opt_float3, cells, 1)
local rest   = Array("UrShape", opt_float3, cells, 2)
local handle = Array("Constraints", opt_float3, cells, 3)
local sqrtHandle, sqrtRigid = Param("w_fitSqrt", float, 4), Param("w_regSqrt", float, 5)
UsePreconditioner(true)

local isHandle = greatereq(handle(0, 0, 0)(0), -999999.9)
Energy(Select(isHandle, sqrtHandle * (pos(0, 0, 0) - handle(0, 0, 0)), 0))

-- six axis neighbours; a term only exists where both lattice points exist
local axes = { { 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 } }
for _, n in ipairs(axes) do
    local i, j, k = n[1], n[2], n[3]
    local restEdge = rest(0, 0, 0) - rest(i, j, k)
    local nowEdge  = pos(0, 0, 0) - pos(i, j, k)
    local mismatch = nowEdge - Rotate3D(euler(0, 0, 0), restEdge)
    Energy(sqrtRigid * Select(InBounds(0, 0, 0), Select(InBounds(i, j, k), mismatch, 0.0), 0.0))
end
