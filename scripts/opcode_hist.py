import sys
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "reference-qvm_b200"))
import numpy as np
from referenceqvm import _engine, gates as G
from oracle.qvm_oracle import random_circuit_spec
n = 28
eng = _engine.Engine(n)
eng.reset()
for name, params, qubits in random_circuit_spec(n, 40, 1234):
    m = G.gate_matrix[name]
    eng.queue_matrix(m(*params) if params else m, qubits)
eng.flush(); eng.sync()
h = eng.state.opcode_hist()
names = {0: "X_RELABEL", 1: "H0", 2: "H1", 3: "H2", 4: "H3", 5: "D1_0", 6: "D1_1", 7: "D1_2", 8: "D1_3", 9: "XS0", 10: "XS1", 11: "XS2", 12: "XS3", 13: "D2_0", 14: "D2_1", 15: "DIAG_T", 16: "DIAG_S", 17: "DIAG1_0", 18: "DIAG1_1", 19: "DIAG1_2", 20: "DIAG1_3", 21: "DIAG_G"}
st = eng.stats()
print({names[i]: int(v) for i, v in enumerate(h) if v}, "groups", st["groups_launched"], "rounds", st["regtile_rounds"])
