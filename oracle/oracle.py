"""Python handle on the CPU oracle (oracle/libllq_cpu.so).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's CPU legs -- never from the product package."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libllq_cpu.so")


def build(force=False):
    src = os.path.join(_HERE, "llq_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return LIB_PATH


def load():
    from lifelike_agility_and_play_b200._capi import LlqLibrary
    if not os.path.exists(LIB_PATH):
        build()
    return LlqLibrary(LIB_PATH)


def make_engine(n_envs, model_blob, mocap, **cfg):
    from lifelike_agility_and_play_b200._capi import VecEngine
    return VecEngine(load(), n_envs, model_blob, mocap, **cfg)
