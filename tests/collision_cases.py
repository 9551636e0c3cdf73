"""Known-answer narrow-phase cases (closed forms) that BOTH the fp64 oracle and the CUDA kernel must reproduce.

Each case: a static geom S (world child) and a free body F carrying one geom, placed by qpos; expected number of
contacts, expected `dist` of every contact, expected normal (geom1 -> geom2; the static geom comes first in the
model, and geom1 is the one with the lower type id, ties by id).  Margin 0.01 (the levels' `<geom margin>`)."""
import math

import numpy as np

GAP = 0.004


def _quat(axis, deg):
    a = np.asarray(axis, float)
    a = a / np.linalg.norm(a)
    h = math.radians(deg) / 2
    return [math.cos(h)] + list(a * math.sin(h))


def _qmul(a, b):
    return [a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
            a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]]


def scene(static_geom, free_geom):
    return f"""<mujoco><default><geom margin="0.01" friction="1 0.5 0.5" density="5"/></default><worldbody>
      {static_geom}
      <body name="F" pos="0 0 5"><freejoint/>{free_geom}</body></worldbody></mujoco>"""


BIGBOX = '<geom name="S" type="box" size="1 1 0.5" pos="0 0 0"/>'
PLANE = '<geom name="S" type="plane" size="10 10 0.1" pos="0 0 0"/>'
CAPS = '<geom name="G" type="capsule" size="0.1 0.3"/>'        # axis = local z
SMALL = '<geom name="G" type="box" size="0.2 0.3 0.1"/>'
s30, c30 = math.sin(math.radians(30)), math.cos(math.radians(30))

# name, xml, qpos of F (pos + quat), expected contact count, expected dists (sorted), expected normal (or None)
CASES = [
    # capsule lying flat on the top face: the two ends of the segment, both at the gap
    ("capsule_flat_on_face", scene(BIGBOX, CAPS), [0.1, -0.2, 0.5 + 0.1 + GAP] + _quat([0, 1, 0], 90), 2, [GAP, GAP], [0, 0, -1]),
    # capsule lying on the face and sticking out over the edge x = 1: inner end + the point above the edge
    ("capsule_overhanging_edge", scene(BIGBOX, CAPS), [1.0, 0.0, 0.5 + 0.1 + GAP] + _quat([0, 1, 0], 90), 2, [GAP, GAP], [0, 0, -1]),
    # capsule overhanging on both sides (box narrower than the capsule): the two points above the edges
    ("capsule_bridging_box", scene('<geom name="S" type="box" size="0.15 1 0.5"/>', CAPS), [0.02, 0.3, 0.5 + 0.1 + GAP] + _quat([0, 1, 0], 90),
     2, [GAP, GAP], [0, 0, -1]),
    # capsule leaning on the edge (x = 1, z = 0.5) with its middle: one contact, normal along the common perpendicular
    ("capsule_leaning_on_edge", scene(BIGBOX, CAPS), [1 + (0.1 + GAP) * s30, 0.1, 0.5 + (0.1 + GAP) * c30] + _quat([0, 1, 0], 120), 1, [GAP],
     [-s30, 0, -c30]),
    # capsule standing upright on the face (T configuration): one end only
    ("capsule_upright", scene(BIGBOX, CAPS), [0.3, 0.3, 0.5 + 0.3 + 0.1 + GAP, 1, 0, 0, 0], 1, [GAP], [0, 0, -1]),
    # capsule end poking into the face 1 cm deep
    ("capsule_penetrating", scene(BIGBOX, CAPS), [0.3, 0.3, 0.5 + 0.3 + 0.1 - 0.01, 1, 0, 0, 0], 1, [-0.01], [0, 0, -1]),
    # box resting on the bigger box: the four corners of its bottom face
    ("box_on_face", scene(BIGBOX, SMALL), [0.1, 0.2, 0.5 + 0.1 + GAP] + _quat([0, 0, 1], 25), 4, [GAP] * 4, [0, 0, 1]),
    # box half over the edge x = 1: two corners inside + two points where its bottom edges cross the big box's edge
    ("box_overhanging", scene(BIGBOX, SMALL), [1.0, 0.0, 0.5 + 0.1 + GAP, 1, 0, 0, 0], 4, [GAP] * 4, [0, 0, 1]),
    # bigger face on a smaller one (incident face covers the reference face): the reference rectangle's four corners
    ("box_covering_pillar", scene('<geom name="S" type="box" size="0.1 0.15 0.5"/>', SMALL), [0.02, 0.03, 0.5 + 0.1 + GAP] + _quat([0, 0, 1], 10),
     4, [GAP] * 4, [0, 0, 1]),
    # rotated 45 deg about z over a pillar of comparable size: octagonal overlap, 8 contacts
    ("box_octagon", scene('<geom name="S" type="box" size="0.2 0.2 0.5"/>', '<geom name="G" type="box" size="0.2 0.2 0.1"/>'),
     [0, 0, 0.5 + 0.1 + GAP] + _quat([0, 0, 1], 45), 8, [GAP] * 8, [0, 0, 1]),
    # tilted 10 deg about y: only the lower bottom edge (two corners) is within the margin
    ("box_tilted_edge_down", scene(BIGBOX, SMALL),
     [0, 0, 0.5 + GAP + 0.2 * math.sin(math.radians(10)) + 0.1 * math.cos(math.radians(10))] + _quat([0, 1, 0], 10), 2, [GAP, GAP], [0, 0, 1]),
    # edge on edge, crossed: static box turned 45 deg about y (top edge along y), free box 45 deg about x (bottom edge along x)
    ("box_edge_on_edge", scene('<geom name="S" type="box" size="0.3 0.3 0.3" euler="0 45 0"/>', '<geom name="G" type="box" size="0.2 0.2 0.2"/>'),
     [0.05, -0.04, 0.3 * math.sqrt(2) + 0.2 * math.sqrt(2) + GAP] + _quat([1, 0, 0], 45), 1, [GAP], [0, 0, 1]),
    # edge on edge penetrating 5 mm
    ("box_edge_on_edge_deep", scene('<geom name="S" type="box" size="0.3 0.3 0.3" euler="0 45 0"/>', '<geom name="G" type="box" size="0.2 0.2 0.2"/>'),
     [0.05, -0.04, 0.3 * math.sqrt(2) + 0.2 * math.sqrt(2) - 0.005] + _quat([1, 0, 0], 45), 1, [-0.005], [0, 0, 1]),
    # corner of a box on the plane: cube standing on its (1,1,1) diagonal
    ("box_corner_on_plane", scene(PLANE, '<geom name="G" type="box" size="0.2 0.2 0.2"/>'),
     [0, 0, 0.2 * math.sqrt(3) + GAP] + _qmul(_quat([1, -1, 0], math.degrees(math.acos(1 / math.sqrt(3)))), [1, 0, 0, 0]), 1, [GAP], [0, 0, 1]),
    # box flat on the plane: four corners
    ("box_flat_on_plane", scene(PLANE, SMALL), [0.3, 0.1, 0.1 + GAP] + _quat([0, 0, 1], 33), 4, [GAP] * 4, [0, 0, 1]),
    # just out of reach: nothing
    ("box_out_of_margin", scene(BIGBOX, SMALL), [0.1, 0.2, 0.5 + 0.1 + 0.011, 1, 0, 0, 0], 0, [], None),
    ("capsule_out_of_margin", scene(BIGBOX, CAPS), [0.1, -0.2, 0.5 + 0.1 + 0.011] + _quat([0, 1, 0], 90), 0, [], None),
]
