"""Writes tests/golden/reference_known_answers.json.

The reference (Julia) cannot be executed in this image, so these vectors are TRANSCRIBED from the reference's
own known-answer tests and documentation, each entry citing the file:line under /root/reference it comes from.
They pin the oracle (tests/test_oracle_golden.py) and, through it and directly, the CUDA library.

Matrices are written row by row exactly as the Julia source prints them: entry [i][j] is out[i+1, j+1].
Points are lists of N_in-vectors; rotations are row-major lists of rows (so [[0,-1],[1,0]] is the Julia
literal Float64[0; 1;; -1; 0], whose columns are (0,1) and (-1,0)).
"""
import json
import os

Z = [0, 0, 0, 0, 0]

single_center = [[0.0, 0.0]]
single_1pix_right = [[0.0, 0.4]]
single_1pix_up = [[-0.4, 0.0]]
single_1pix_left = [[0.0, -0.4]]
single_1pix_down = [[0.4, 0.0]]
single_halfpix_down = [[0.2, 0.0]]
single_halfpix_down_and_right = [[0.2, 0.2]]
four_cross = single_1pix_right + single_1pix_up + single_1pix_left + single_1pix_down   # src/raster.jl:121-129

no_rotation = [[1.0, 0.0], [0.0, 1.0]]            # src/raster.jl:131
rotation_90_deg = [[0.0, -1.0], [1.0, 0.0]]       # src/raster.jl:132
no_translation = [0.0, 0.0]
translation_halfpix_right = [0.0, 0.2]            # src/raster.jl:135
translation_1pix_down = [0.4, 0.0]                # src/raster.jl:136


def fwd(name, cite, points, rotation, translation, background, out_weight, point_weight, expected):
    return dict(name=name, cite=cite, grid_size=[5, 5], points=points, rotation=rotation, translation=translation,
                background=background, out_weight=out_weight, point_weight=point_weight, expected=expected)


forward = [
    fwd("single_center", "src/raster.jl:143-157", single_center, no_rotation, no_translation, 0.0, 4.0, None,
        [Z, Z, [0, 0, 4, 0, 0], Z, Z]),
    fwd("single_1pix_right", "src/raster.jl:159-173", single_1pix_right, no_rotation, no_translation, 0.0, 4.0, None,
        [Z, Z, [0, 0, 0, 4, 0], Z, Z]),
    fwd("single_halfpix_down", "src/raster.jl:175-189", single_halfpix_down, no_rotation, no_translation, 0.0, 4.0, None,
        [Z, Z, [0, 0, 2, 0, 0], [0, 0, 2, 0, 0], Z]),
    fwd("single_halfpix_down_and_right", "src/raster.jl:191-205", single_halfpix_down_and_right, no_rotation,
        no_translation, 0.0, 4.0, None, [Z, Z, [0, 0, 1, 1, 0], [0, 0, 1, 1, 0], Z]),
    fwd("four_cross", "src/raster.jl:209-223", four_cross, no_rotation, no_translation, 0.0, 4.0, None,
        [Z, [0, 0, 4, 0, 0], [0, 4, 0, 4, 0], [0, 0, 4, 0, 0], Z]),
    fwd("four_cross_translation_halfpix_right", "src/raster.jl:225-239", four_cross, no_rotation,
        translation_halfpix_right, 0.0, 4.0, None, [Z, [0, 0, 2, 2, 0], [0, 2, 2, 2, 2], [0, 0, 2, 2, 0], Z]),
    fwd("four_cross_translation_1pix_down", "src/raster.jl:241-255", four_cross, no_rotation, translation_1pix_down,
        0.0, 4.0, None, [Z, Z, [0, 0, 4, 0, 0], [0, 4, 0, 4, 0], [0, 0, 4, 0, 0]]),
    fwd("single_1pix_right_rotation_90", "src/raster.jl:259-273", single_1pix_right, rotation_90_deg, no_translation,
        0.0, 4.0, None, [Z, [0, 0, 4, 0, 0], Z, Z, Z]),
    fwd("four_cross_point_weights", "src/raster.jl:277-292", four_cross, no_rotation, no_translation, 0.0, 1.0,
        [1.0, 2.0, 3.0, 4.0], [Z, [0, 0, 2, 0, 0], [0, 3, 0, 1, 0], [0, 0, 4, 0, 0], Z]),
    fwd("four_cross_point_weights_out_weight_translation", "src/raster.jl:294-309", four_cross, no_rotation,
        translation_halfpix_right, 0.0, 2.0, [1.0, 2.0, 3.0, 4.0],
        [Z, [0, 0, 2, 2, 0], [0, 3, 3, 1, 1], [0, 0, 4, 4, 0], Z]),
    # README examples use the default background (0) and weights (1)
    fwd("readme_center", "README.md:41-48", [[0.0, 0.0]], no_rotation, no_translation, None, None, None,
        [Z, Z, [0, 0, 1.0, 0, 0], Z, Z]),
    fwd("readme_halfpix_below", "README.md:51-58", [[0.2, 0.0]], no_rotation, no_translation, None, None, None,
        [Z, Z, [0, 0, 0.5, 0, 0], [0, 0, 0.5, 0, 0], Z]),
    fwd("readme_halfpix_below_left", "README.md:61-68", [[0.2, -0.2]], no_rotation, no_translation, None, None, None,
        [Z, Z, [0, 0.25, 0.25, 0, 0], [0, 0.25, 0.25, 0, 0], Z]),
]
# note: src/raster.jl:143-309 holds ten `@test out ≈ [...]` blocks (SURVEY.md says eleven; `grep -c "@test out"`
# over src/raster.jl:110-310 gives 10) - all ten are listed above, followed by the three README matrices.

pullback = dict(
    cite="README.md:84-90 (target_image), :99-105 (points), :107-113 (rotation), :36 (translation), "
         ":151-157 (ds_dout), :169-182 (explicit pullback), :120-137 (Zygote gradient = minus the pullback)",
    grid_size=[5, 5],
    points=[[0.8457397177007744, 0.3482756109584688],
            [-0.6028188536164718, -0.612801322279686],
            [-0.47141692007256464, 0.6098964840013308],
            [-0.74526926786903, 0.6480225109030409],
            [-0.4044384373422192, -0.13171854413805173]],
    rotation=[[1.0, 0.0], [0.0, 1.0]],
    translation=[0.0, 0.0],
    target_image=[[0.345889, 0.032283, 0.589178, 0.0625972, 0.310929],
                  [0.404836, 0.573265, 0.350633, 0.0417926, 0.895955],
                  [0.174528, 0.127555, 0.0906833, 0.639844, 0.832502],
                  [0.189836, 0.360597, 0.243664, 0.825484, 0.667319],
                  [0.672631, 0.520593, 0.341701, 0.101026, 0.182172]],
    # printed with 6 significant digits, so everything derived from it is good to ~1e-5
    ds_dout=[[0.152276, -0.417335, 1.16347, -0.700428, -0.63595],
             [0.285167, 0.033845, -0.625258, -0.801198, 0.760124],
             [0.349055, 0.25511, 0.181367, 1.27969, 1.665],
             [0.379672, 0.721194, 0.487329, 1.65097, 1.33464],
             [1.34526, 1.04119, 0.454354, -1.3402, 0.364343]],
    # README.md:169-172: 2x5 matrix, column p = gradient of point p
    d_points=[[2.77036, 0.704622, 1.71171, 2.06839, -2.62784],
              [-3.97337, -1.03177, 3.23518, 0.673215, -1.58562]],
    d_rotation=[[0.632605, 3.26353], [-4.12402, 1.86668]],          # README.md:174-177
    d_translation=[4.62725350082958, -2.6823723442258274],           # README.md:179-182 (full digits)
    # README.md:120-126 full-precision Zygote gradient of the loss = MINUS the pullback of ds_dout
    zygote_d_points=[[-2.7703628931165025, 3.973371400200988],
                     [-0.70462225282373, 1.0317734946448016],
                     [-1.7117138793471494, -3.235178706903591],
                     [-2.0683933141077886, -0.6732149105779637],
                     [2.6278388385655904, 1.585621066861592]],
)

voxel_shifts = {   # src/util.jl:32-45
    "1": [[0], [1]],
    "2": [[0, 0], [1, 0], [0, 1], [1, 1]],
    "3": [[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]],
}

if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "reference_known_answers.json"), "w") as f:
        json.dump(dict(forward=forward, pullback=pullback, voxel_shifts=voxel_shifts), f, indent=1)
    print("wrote", len(forward), "forward cases")
