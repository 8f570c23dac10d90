"""Limb-tree fixtures (pre-order ``parents`` lists) of the morphology sets the reference
ships under ``src/environments/3d_*`` (derived with the pre-order rule of
src/utils.py:236-276; listed in SURVEY.md Appendix C).  MuJoCo XML parsing itself is
host glue and out of scope; these lists are all the SET hot path needs from it."""
from __future__ import annotations

HOPPERS = {
    "3d_hopper_3_shin": [-1, 0, 1],
    "3d_hopper_4_lower_shin": [-1, 0, 1, 2],
    "3d_hopper_5_full": [-1, 0, 1, 2, 3],
}

WALKERS = {
    "3d_walker_2_right_leg_left_knee": [-1, 0],
    "3d_walker_3_left_leg_right_foot": [-1, 0, 1],
    "3d_walker_4_right_knee_left_foot": [-1, 0, 0, 2],
    "3d_walker_5_foot": [-1, 0, 1, 0, 3],
    "3d_walker_5_left_knee": [-1, 0, 1, 2, 0],
    "3d_walker_7_full": [-1, 0, 1, 2, 0, 4, 5],
}
WALKERS_HELD_OUT = {
    "3d_walker_3_left_knee_right_knee": [-1, 0, 0],
    "3d_walker_6_right_foot": [-1, 0, 1, 0, 3, 4],
}

HUMANOIDS = {
    "3d_humanoid_7_left_arm": [-1, 0, 1, 0, 3, 0, 5],
    "3d_humanoid_7_lower_arms": [-1, 0, 1, 0, 3, 0, 0],
    "3d_humanoid_7_right_arm": [-1, 0, 1, 0, 3, 0, 5],
    "3d_humanoid_7_right_leg": [-1, 0, 1, 0, 3, 0, 5],
    "3d_humanoid_8_left_knee": [-1, 0, 1, 0, 0, 4, 0, 6],
    "3d_humanoid_9_full": [-1, 0, 1, 0, 3, 0, 5, 0, 7],
}
HUMANOIDS_HELD_OUT = {
    "3d_humanoid_7_left_leg": [-1, 0, 1, 0, 3, 0, 5],
    "3d_humanoid_8_right_knee": [-1, 0, 0, 2, 0, 4, 0, 6],
}

CHEETAHS = {
    "3d_cheetah_10_tail_leftbleg": [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8],
    "3d_cheetah_11_leftfleg": [-1, 0, 0, 2, 3, 0, 5, 6, 0, 8, 9],
    "3d_cheetah_11_tail_rightfknee": [-1, 0, 1, 2, 0, 0, 5, 6, 0, 8, 9],
    "3d_cheetah_12_rightbknee": [-1, 0, 0, 0, 3, 4, 0, 6, 7, 0, 9, 10],
    "3d_cheetah_12_tail_leftbfoot": [-1, 0, 1, 2, 0, 4, 5, 0, 7, 0, 9, 10],
    "3d_cheetah_13_rightffoot": [-1, 0, 0, 2, 3, 0, 5, 0, 7, 8, 0, 10, 11],
    "3d_cheetah_13_tail": [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11],
    "3d_cheetah_14_full": [-1, 0, 0, 2, 3, 0, 5, 6, 0, 8, 9, 0, 11, 12],
}

CWHH = {**CHEETAHS, **HOPPERS, **HUMANOIDS, **WALKERS}

SETS = {
    "3d_hoppers": HOPPERS,
    "3d_walkers": WALKERS,
    "3d_humanoids": HUMANOIDS,
    "3d_cheetahs": CHEETAHS,
    "3d_cwhh": CWHH,
}

ALL = {**CWHH, **WALKERS_HELD_OUT, **HUMANOIDS_HELD_OUT}

MAX_LIMBS = 15  # rows of the positional-embedding tables (src/SEActor.py:19)
