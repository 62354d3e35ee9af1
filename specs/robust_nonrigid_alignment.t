-- Non-rigid alignment of a template mesh to scanned points with per-vertex robustness weights (test specification of
-- opt_b200; the parameter table matches the robust_nonrigid_alignment example of niessner/Opt).
--   slot 0 / 1: square roots of the data and rigidity weights
--   slots 2..4, UNKNOWN: vertex positions, vertex Euler angles, per-vertex robustness weight
--   slot 5: rest positions      slot 6: correspondences (below -999999.9 = none)      slot 7: correspondence normals
--   slot 8: number of directed edges      slots 9, 10: edge tail / head
local verts = { Dim("N", 0) }
local sqrtData, sqrtRigid = Param("w_fitSqrt", float, 0), Param("w_regSqrt", float, 1)
local sqrtConfidence = 0.1
local pos    = Unknown("Offset", opt_float3, verts, 2)
local euler  = Unknown("Angle", opt_float3, verts, 3)
local trust  = Unknown("RobustWeights", opt_float, verts, 4)
local rest   = Array("UrShape", opt_float3, verts, 5)
local match  = Array("Constraints", opt_float3, verts, 6)
local normal = Array("ConstraintNormals", opt_float3, verts, 7)
local edges  = Graph("G", 8, "v0", verts, 9, "v1", verts, 10)
UsePreconditioner(true)

local t = trust(0)
-- NOTE: this is a test of a 3-vector, channel by channel: each of the two terms below exists once per valid channel
local matched = greatereq(match(0), -999999.9)

-- signed distance to the tangent plane of the correspondence, damped by the robustness weight
Energy(sqrtData * Select(matched, t * normal(0):dot(pos(0) - match(0)), 0.0))
-- the robustness weight should not drift away from one without a reason
Energy(sqrtConfidence * Select(matched, 1 - (t * t), 0.0))

-- rigidity along the mesh edges
local tail, head = edges.v0, edges.v1
Energy(sqrtRigid * ((pos(tail) - pos(head)) - Rotate3D(euler(tail), rest(tail) - rest(head))))
