-- Intrinsic images: split a colour image (in the log domain) into reflectance and a scalar shading layer, with a sparse
-- (L_p) prior on reflectance changes (test specification of opt_b200; the parameter table matches the
-- intrinsic_image_decomposition example of niessner/Opt).  No built-in kernel set: general front end.
--   slots 0..2: square roots of the data, reflectance-prior and shading-prior weights      slot 3: exponent p (opt_float)
--   slot 4: reflectance, 3 channels, UNKNOWN -- the same buffer is also read as constant data to build the L_p weights
--   slot 5: input image      slot 6: shading, UNKNOWN
local pix = { Dim("W", 0), Dim("H", 1) }
local sqrtData, sqrtSparse, sqrtSmooth = Param("w_fitSqrt", float, 0), Param("w_regSqrtAlbedo", float, 1), Param("w_regSqrtShading", float, 2)
local exponent = Param("pNorm", opt_float, 3)
local albedo       = Unknown("r", opt_float3, pix, 4)
local albedoFrozen = Array("r_const", opt_float3, pix, 4)
local picture      = Array("i", opt_float3, pix, 5)
local shade        = Unknown("s", opt_float, pix, 6)

local function bothInside(dx, dy, term) return Select(InBounds(0, 0), Select(InBounds(dx, dy), term, 0), 0) end
local cross = { { 1, 0 }, { -1, 0 }, { 0, 1 }, { 0, -1 } }
-- reflectance changes are sparse: iteratively re-weighted with the frozen copy
for _, d in ipairs(cross) do
    local change = albedo(0, 0) - albedo(d[1], d[2])
    local changeFrozen = albedoFrozen(0, 0) - albedoFrozen(d[1], d[2])
    Energy(sqrtSparse * bothInside(d[1], d[2], L_p(change, changeFrozen, exponent, pix)))
end
-- shading varies smoothly
for _, d in ipairs(cross) do
    Energy(sqrtSmooth * bothInside(d[1], d[2], shade(0, 0) - shade(d[1], d[2])))
end
-- the two layers add up to the picture
Energy(sqrtData * (albedo(0, 0) + shade(0, 0) - picture(0, 0)))
