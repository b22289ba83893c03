import sys, numpy as np
sys.path.insert(0, ".")
from orb_slam3_study_kr_b200 import api, synthetic, problem
from orb_slam3_study_kr_b200.problem import Round, Schedule, GATE_LBA, GATE_NONE
from oracle import ba_ref
ctx = api.Context(0)
p = synthetic.config(1, scale=0.2)
def run(name, s):
    got = ctx.solve_ba(p, s); ref = ba_ref.solve(p, s)
    w = max(abs(a['chi2_after']-b['chi2_after'])/b['chi2_after'] for a,b in zip(got.trace, ref.trace))
    print(name, "worst rel", w, [t['trials'] for t in got.trace], [t['trials'] for t in ref.trace])
run("gate+drop", Schedule([Round(5, GATE_LBA, 5.991, 7.815, True), Round(10)], problem.DELTA_MONO_GBA, problem.DELTA_STEREO))
run("gate only", Schedule([Round(5, GATE_LBA, 5.991, 7.815, False), Round(10)], problem.DELTA_MONO_GBA, problem.DELTA_STEREO))
run("drop only", Schedule([Round(5, GATE_NONE, 5.991, 7.815, True), Round(10)], problem.DELTA_MONO_GBA, problem.DELTA_STEREO))
run("two rounds plain", Schedule([Round(5), Round(10)], problem.DELTA_MONO_GBA, problem.DELTA_STEREO))
