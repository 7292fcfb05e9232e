"""Gate opcodes of the C ABI (include/bwq.h) and their Qiskit names.

Names follow the reference's list of supported instructions, blackwater/data/utils.py:19-49,
plus the backend basis {id, rz, sx, x, cx, reset} (docs/tutorials/02_data_generation.ipynb cell 3).
"""

OPCODES = {
    "id": 0, "x": 1, "y": 2, "z": 3, "h": 4, "s": 5, "sdg": 6, "t": 7, "tdg": 8, "sx": 9, "sxdg": 10,
    "rx": 11, "ry": 12, "rz": 13, "p": 14, "u2": 15, "u3": 16, "reset": 17,
    "cx": 32, "cy": 33, "cz": 34, "ch": 35, "crx": 36, "cry": 37, "crz": 38, "cp": 39, "cu3": 40,
    "swap": 41, "iswap": 42, "rzz": 43, "rxx": 44, "ryy": 45, "rzx": 46, "ecr": 47,
    "unitary1": 64, "unitary2": 65,
}
ALIASES = {"i": "id", "u1": "p", "u": "u3", "cnot": "cx", "cu1": "cp", "cphase": "cp"}
NUM_PARAMS = {
    "rx": 1, "ry": 1, "rz": 1, "p": 1, "u2": 2, "u3": 3, "crx": 1, "cry": 1, "crz": 1, "cp": 1, "cu3": 3,
    "rzz": 1, "rxx": 1, "ryy": 1, "rzx": 1, "unitary1": 8, "unitary2": 32,
}
IGNORED = ("barrier", "delay", "snapshot")
NAMES = {v: k for k, v in OPCODES.items()}


def canonical(name):
    n = name.lower()
    return ALIASES.get(n, n)


def is_two_qubit(name):
    return 32 <= OPCODES[name] <= 47 or name == "unitary2"
