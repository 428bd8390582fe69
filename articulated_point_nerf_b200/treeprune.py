"""Bone pruning and sibling merging on the kinematic tree (host-side index logic, CPU / numpy).

Same contract as the reference's ``merge_joints`` (lib/treeprune.py:41-228, helper ``cluster_children`` :5-39),
called by ``TemporalPoints.simplify_skeleton`` (lib/temporalpoints.py:256-343): given the joints, the bone list
(parent, child), a per-joint "zero motion" mask and a joint-by-joint "rotations are similar" matrix it returns

    new_joints, new_bones, merging_rules, joints_to_keep, rotations_to_keep, rotation_switch_mask,
    sibling_transfer_rules

Only ``merging_rules`` and ``sibling_transfer_rules`` reach the render path (they become ``flat_merging_rules``, the
skinning-weight column merge of the LBS kernel, and ``forward_warp.sibling_mask``); the other outputs describe the
pruned tree for visualisation.  The reference's observable quirks are kept on purpose, each
marked "quirk" below; tests/test_cpu_host.py pins every output against the reference run on its own 29-joint
fixture and on random trees (tests/golden/ref_treeprune.pt).
"""
from __future__ import annotations

from itertools import combinations

import numpy as np


class _Tree:
    """parent / children tables of a bone list [(parent, child), ...]; children keep the bone-list order."""

    def __init__(self, n_joints, bones, root):
        self.n = n_joints
        self.root = root
        self.parent = {}
        self.children = [[] for _ in range(n_joints)]
        for p, c in bones:
            self.parent[int(c)] = int(p)
        for c, p in self.parent.items():          # dict order = first appearance as a bone tail
            self.children[p].append(c)
        self.leaves = [j for j in range(n_joints) if not self.children[j]]
        self.branching = [len(ch) > 1 for ch in self.children]

    def walk_up(self, j):
        """j, parent(j), ... up to but excluding the root."""
        while j != self.root:
            yield j
            j = self.parent[j]


def _similar_sibling_groups(siblings, similar):
    """Transitive-ish grouping of siblings whose motion is similar -> {kept sibling: array of absorbed siblings}.

    quirks kept: a pair joins EVERY existing group that already holds one of its members (groups are never fused);
    the kept sibling is the first element in the iteration order of a Python ``set`` (lib/treeprune.py:17-37)."""
    groups = []
    for a, b in combinations(siblings, 2):
        if not similar[a, b]:
            continue
        hit = False
        for g in groups:
            if a in g or b in g:
                g.add(a)
                g.add(b)
                hit = True
        if not hit:
            groups.append(set((a, b)))
    out = {}
    for g in groups:
        members = np.array(list(g))
        out[members[0]] = members[1:]
    return out


def _pruned_paths(tree, prune):
    """For every leaf: the root->leaf path with pruned joints skipped, and the full root->leaf path.

    A joint contributes its PARENT to the pruned path when it moves (not pruned) or hangs off a branching joint;
    the lowest such joint is itself included only if its parent does not branch (quirk, lib/treeprune.py:66-70:
    a moving leaf directly under a branching joint is dropped from the new tree)."""
    kept_paths, full_paths = [], []
    for leaf in tree.leaves:
        kept, full = [], []
        for j in tree.walk_up(leaf):
            p = tree.parent[j]
            if (not prune[j]) or tree.branching[p]:
                if not kept and not tree.branching[p]:
                    kept.append(j)
                kept.append(p)
            full.append(j)
        if not kept or kept[-1] != tree.root:
            kept.append(tree.root)
        full.append(tree.root)
        kept_paths.append(kept[::-1])
        full_paths.append(full[::-1])
    return kept_paths, full_paths


def merge_joints(joints, bones, prune_bones, rotation_similarity_matrix, root_idx=0, convert_merging_rules=True):
    joints = np.asarray(joints)
    prune = np.asarray(prune_bones).astype(bool)
    assert len(joints) == len(prune)
    n = len(joints)
    tree = _Tree(n, bones, root_idx)
    kept_paths, full_paths = _pruned_paths(tree, prune)

    # ---- bones of the pruned tree: consecutive pairs of the pruned paths (a set: its iteration order fixes the
    # order of the intermediate arrays exactly as in the reference, lib/treeprune.py:87-95)
    edge_set = set()
    for path in kept_paths:
        for a, b in zip(path[:-1], path[1:]):
            edge_set.add((a, b))
    edges = np.array([[a, b] for a, b in edge_set])
    kept_idx = np.unique(edges)
    new_joints = joints[kept_idx]

    # ---- which original rotation drives each new bone: the child of the bone's head on the way to its tail
    # (quirk, lib/treeprune.py:103-116: with several children and no full path through child and tail the LAST
    # child is used)
    drivers = []
    for head, tail in edges:
        ch = tree.children[head]
        pick = ch[-1]
        if len(ch) > 1:
            for c in ch:
                if any((c in path) and (tail in path) for path in full_paths):
                    pick = c
                    break
        else:
            pick = ch[0]
        drivers.append(pick)
    rotations_to_keep = np.zeros(n, dtype=bool)
    rotations_to_keep[drivers] = True
    rotations_to_keep[root_idx] = True

    by_tail = np.argsort(edges[:, 1], axis=0)
    drivers = np.array(drivers)[by_tail]
    # dense renumbering of the drivers (ascending), shifted by one for the root slot
    rank = {old: r for r, old in enumerate(np.unique(drivers))}
    rotation_switch_mask = np.concatenate([[0], np.array([rank[d] for d in drivers], dtype=drivers.dtype) + 1])

    joints_to_keep = np.zeros(n, dtype=bool)
    joints_to_keep[kept_idx] = True

    # ---- new bone list in the new numbering, sorted by tail
    renumber = np.full(n, -1, dtype=edges.dtype)
    renumber[kept_idx] = np.arange(len(kept_idx))
    new_bones = renumber[edges]
    new_bones = new_bones[np.argsort(new_bones[:, 1], axis=0)]

    # ---- weight merging rules in the ORIGINAL numbering: a pruned joint hands its weight to its nearest proper
    # ancestor that moves, or to the root (lib/treeprune.py:155-181)
    merging_rules = np.arange(n, dtype=np.int16)
    for leaf in tree.leaves:
        waiting = []
        for j in tree.walk_up(leaf):
            if prune[j]:
                waiting.append(j)
            else:
                merging_rules[waiting] = j
                waiting = []
        merging_rules[waiting] = root_idx

    # ---- siblings that were not merged upwards and move alike share one rotation (lib/treeprune.py:188-201)
    sibling_transfer_rules = np.arange(n, dtype=np.int16)
    for ch in tree.children:
        free = [c for c in ch if merging_rules[c] == c]
        if len(free) > 1:
            for keep, absorbed in _similar_sibling_groups(free, rotation_similarity_matrix).items():
                merging_rules[absorbed] = keep
                sibling_transfer_rules[absorbed] = keep

    if convert_merging_rules:
        # express the rule TARGETS through joints that survive in the pruned paths (lib/treeprune.py:204-226)
        survivor = {}
        for kept, full in zip(kept_paths, full_paths):
            waiting = []
            for j in full:
                if j in kept:
                    for w in waiting:
                        survivor[w] = j
                    survivor[j] = j
                    waiting = []
                else:
                    waiting.append(j)
        converted = merging_rules.copy()
        for old in range(n):
            if old in survivor:
                converted[merging_rules == old] = survivor[old]
        merging_rules = converted

    return new_joints, new_bones, merging_rules, joints_to_keep, rotations_to_keep, rotation_switch_mask, \
        sibling_transfer_rules
