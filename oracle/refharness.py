"""TEST INFRASTRUCTURE ONLY.  ctypes view of oracle/_ref/liboxref.so (the unmodified reference CPU
implementation behind our C harness oracle/ref_harness.cpp).  Never imported by the product package.

Only usable where oracle/_ref/ has been built (this container; the .so also travels to the GPU box).
"""
import ctypes as C
import os
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liboxref.so")


def available():
    return os.path.exists(_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_SO, )
        _lib.oxref_last_error.restype = C.c_char_p
        _lib.oxref_rcut.restype = C.c_double
        _lib.oxref_temperature.restype = C.c_double
        _lib.oxref_compute_forces.restype = C.c_double
        _lib.oxref_system_energy.restype = C.c_double
        _lib.oxref_get_pairs.restype = C.c_longlong
        _lib.oxref_current_step.restype = C.c_longlong
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


DEFAULT_INPUT = """
backend = CPU
sim_type = MD
steps = 0
dt = 0.003
thermostat = no
verlet_skin = 0.05
refresh_vel = 0
restart_step_counter = 1
time_scale = linear
print_conf_interval = 100000000
print_energy_every = 100000000
no_stdout_energy = 1
log_file = /dev/null
trajectory_file = /dev/null
lastconf_file = /dev/null
energy_file = /dev/null
"""


class Reference:
    """One reference CPU simulation (a process-wide singleton in the reference: one at a time)."""

    def __init__(self, topology, conf, **keys):
        self.tmp = tempfile.TemporaryDirectory()
        inp = os.path.join(self.tmp.name, "input")
        text = DEFAULT_INPUT + f"topology = {topology}\nconf_file = {conf}\n"
        over = "".join(f"{k} = {v}\n" for k, v in keys.items())
        with open(inp, "w") as f:
            f.write(text)
        L = lib()
        cwd = os.getcwd()
        os.chdir(self.tmp.name)
        try:
            rc = L.oxref_open(inp.encode(), over.encode())
        finally:
            os.chdir(cwd)
        if rc != 0:
            raise RuntimeError("reference: " + L.oxref_last_error().decode())
        self.N = L.oxref_N()

    def close(self):
        lib().oxref_close()
        self.tmp.cleanup()

    # ---- topology / state
    def topology(self):
        a = [np.zeros(self.N, dtype=np.int32) for _ in range(5)]
        lib().oxref_get_topology(*[_p(x) for x in a])
        return dict(btype=a[0], type=a[1], n3=a[2], n5=a[3], strand=a[4])

    def box(self):
        s = np.zeros(3)
        lib().oxref_box(_p(s))
        return s

    def rcut(self):
        return lib().oxref_rcut()

    def state(self):
        a = [np.zeros((self.N, 3)) for _ in range(5)]
        lib().oxref_get_state(*[_p(x) for x in a])
        return dict(pos=a[0], a1=a[1], a3=a[2], vel=a[3], L=a[4])

    def set_state(self, pos, a1, a3, vel=None, L=None):
        c = lambda x: None if x is None else np.ascontiguousarray(x, dtype=np.float64)
        pos, a1, a3, vel, L = c(pos), c(a1), c(a3), c(vel), c(L)
        lib().oxref_set_state(_p(pos), _p(a1), _p(a3), _p(vel), _p(L))

    # ---- physics
    def compute_forces(self):
        U = lib().oxref_compute_forces()
        f, tb, tl = (np.zeros((self.N, 3)) for _ in range(3))
        lib().oxref_get_forces(_p(f), _p(tb), _p(tl))
        return dict(U=U, force=f, torque_body=tb, torque_lab=tl)

    def forces(self):
        f, tb, tl = (np.zeros((self.N, 3)) for _ in range(3))
        lib().oxref_get_forces(_p(f), _p(tb), _p(tl))
        return dict(force=f, torque_body=tb, torque_lab=tl)

    def energy_split(self):
        out = np.zeros(16)
        n = lib().oxref_energy_split(_p(out), 16)
        return out[:n]

    def system_energy(self):
        return lib().oxref_system_energy()

    def pairs(self):
        n = lib().oxref_get_pairs(None, C.c_longlong(0))
        out = np.zeros((max(n, 1), 2), dtype=np.int32)
        lib().oxref_get_pairs(_p(out), C.c_longlong(n))
        return out[:n]

    def step(self, n=1):
        rc = lib().oxref_step(C.c_longlong(n))
        if rc != 0:
            raise RuntimeError("reference: " + lib().oxref_last_error().decode())

    def n_updates(self):
        return lib().oxref_N_updates()

    def update_temperature(self, T):
        lib().oxref_update_temperature(C.c_double(T))
