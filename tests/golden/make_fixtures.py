"""Generates tests/golden/* from the reference checkout (run in the build container only).

  rows_2048x128.f32 : first 2048 rows of rs/index/resources/10000_rows_128_dim (LE f32, uniform[0,1));
                      the reference's own committed dataset for BASELINE config 1.
"""
import os
import numpy as np

REF = "/root/reference/rs/index/resources/10000_rows_128_dim"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    a = np.fromfile(REF, dtype="<f4").reshape(10000, 128)
    a[:2048].tofile(os.path.join(HERE, "rows_2048x128.f32"))
    print("wrote rows_2048x128.f32", a[:2048].shape, float(a.min()), float(a.max()))
