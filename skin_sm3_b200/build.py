"""``python -m skin_sm3_b200.build`` -- compile libsm3_b200.so for sm_100a in-tree."""
from ._lib import build

if __name__ == "__main__":
    print(build(verbose=True))
