-- intrinsic image decomposition with an L_p albedo prior (opt_b200 test spec; same energy and parameter positions as
-- the intrinsic_image_decomposition example of niessner/Opt).  Served by the general front end.
--   params: 0 w_fitSqrt  1 w_regSqrtAlbedo  2 w_regSqrtShading  3 pNorm (opt_float)  4 r (unknown; also bound as the constant view r_const)
--           5 i  6 s (unknown)
local W, H = Dim("W", 0), Dim("H", 1)
local dom = { W, H }
local w_fitSqrt        = Param("w_fitSqrt", float, 0)
local w_regSqrtAlbedo  = Param("w_regSqrtAlbedo", float, 1)
local w_regSqrtShading = Param("w_regSqrtShading", float, 2)
local pNorm            = Param("pNorm", opt_float, 3)
local r       = Unknown("r", opt_float3, dom, 4)
local r_const = Array("r_const", opt_float3, dom, 4)
local i       = Array("i", opt_float3, dom, 5)
local s       = Unknown("s", opt_float, dom, 6)

local neighbours = { { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 } }
for dx, dy in Stencil(neighbours) do
    local step = r(0, 0) - r(dx, dy)
    local frozen = r_const(0, 0) - r_const(dx, dy)
    local sparse = L_p(step, frozen, pNorm, dom)
    Energy(w_regSqrtAlbedo * Select(InBounds(0, 0), Select(InBounds(dx, dy), sparse, 0), 0))
end
for dx, dy in Stencil(neighbours) do
    Energy(w_regSqrtShading * Select(InBounds(0, 0), Select(InBounds(dx, dy), s(0, 0) - s(dx, dy), 0), 0))
end
Energy(w_fitSqrt * (r(0, 0) + s(0, 0) - i(0, 0)))
