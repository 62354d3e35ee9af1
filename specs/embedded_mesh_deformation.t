-- Embedded deformation: every graph node carries a translation and a full 3x3 matrix that is softly kept a rotation
-- (test specification of opt_b200; the parameter table matches the embedded_mesh_deformation example of niessner/Opt).
-- No built-in kernel set: general front end.
--   slots 0..2: square roots of the fit, smoothness and rotation weights
--   slot 3: node positions (UNKNOWN)   slot 4: node matrices, 9 channels row-major (UNKNOWN)
--   slot 5: rest positions   slot 6: handle targets (x below -999999.9 = free node)   slot 7: edge count   slots 8, 9: edge endpoints
local nodes = { Dim("N", 0) }
local sFit, sSmooth, sRot = Param("w_fitSqrt", float, 0), Param("w_regSqrt", float, 1), Param("w_rotSqrt", float, 2)
local place  = Unknown("Offset", opt_float3, nodes, 3)
local affine = Unknown("RotMatrix", opt_float9, nodes, 4)
local rest   = Image("UrShape", opt_float3, nodes, 5)
local target = Image("Constraints", opt_float3, nodes, 6)
local links  = Graph("G", 7, "v0", nodes, 8, "v1", nodes, 9)
UsePreconditioner(true)

-- handles
local isHandle = greatereq(target(0)(0), -999999.9)
Energy(Select(isHandle, sFit * (place(0) - target(0)), 0))

-- orthonormality of the matrix columns: three mixed products vanish, three squared lengths equal one
local m = affine(0)
local function column(j) return Vector(m(j), m(j + 3), m(j + 6)) end
local c0, c1, c2 = column(0), column(1), column(2)
for _, pair in ipairs({ { c0, c1 }, { c0, c2 }, { c1, c2 } }) do
    Energy(sRot * Dot3(pair[1], pair[2]))
end
for _, c in ipairs({ c0, c1, c2 }) do
    Energy(sRot * (Dot3(c, c) - 1))
end

-- a node's matrix should carry its rest edges onto the deformed edges
local from, to = links.v0, links.v1
local restEdge, nowEdge = rest(to) - rest(from), place(to) - place(from)
Energy(sSmooth * (nowEdge - Matrix3x3Mul(affine(from), restEdge)))
