-- Smallest image problem: pull X towards A and flatten its forward differences (test specification of opt_b200; parameter
-- table of tests/minimal in niessner/Opt: slot 0 = X, UNKNOWN, slot 1 = A, both single-channel float images).
local image = { Dim("W", 0), Dim("H", 1) }
local X, A = Unknown("X", float, image, 0), Array("A", float, image, 1)
local dataWeight = 0.2
Energy(dataWeight * (X(0, 0) - A(0, 0)))
-- forward differences; where the neighbour is outside the image the default boundary rule drops the term
for _, step in ipairs({ { 1, 0 }, { 0, 1 } }) do
    Energy(X(0, 0) - X(step[1], step[2]))
end
