-- As-rigid-as-possible deformation of a triangle mesh given as a vertex graph (test specification of opt_b200; the parameter
-- table matches the arap_mesh_deformation example of niessner/Opt).
--   slot 0 / 1: square roots of the handle and rigidity weights
--   slot 2: vertex positions, UNKNOWN      slot 3: vertex Euler angles, UNKNOWN      slot 4: rest positions
--   slot 5: handle targets (x below -999999.9 = free vertex)      slot 6: number of directed edges      slots 7, 8: edge tail / head
local verts = { Dim("N", 0) }
local sqrtHandle, sqrtRigid = Param("w_fitSqrt", float, 0), Param("w_regSqrt", float, 1)
local pos    = Unknown("Offset", opt_float3, verts, 2)
local euler  = Unknown("Angle", opt_float3, verts, 3)
local rest   = Array("UrShape", opt_float3, verts, 4)
local handle = Array("Constraints", opt_float3, verts, 5)
local edges  = Graph("G", 6, "v0", verts, 7, "v1", verts, 8)
UsePreconditioner(true)

-- handle vertices are pulled to their targets
local grabbed = greatereq(handle(0, 0), -999999.9)
Energy(Select(grabbed, sqrtHandle * (pos(0) - handle(0)), 0))

-- along every directed edge the deformed edge should be the rest edge turned by the tail vertex's rotation
local tail, head = edges.v0, edges.v1
local turned = Rotate3D(euler(tail), rest(tail) - rest(head))
Energy(sqrtRigid * ((pos(tail) - pos(head)) - turned))
