"""<base>_probs.h5 (int32 dataset "nlogprobs", H5Segmentation.cpp:25-49) -> <base>_probs.npy for apps/cityscapes_runner.
  python tools/h5_to_npy.py <dir>/probs        (needs h5py, which this image does not have; run it where the data is)"""
import glob, os, sys
import numpy as np

def main(folder):
    import h5py
    for path in sorted(glob.glob(os.path.join(folder, "*_probs.h5"))):
        with h5py.File(path, "r") as f:
            a = np.ascontiguousarray(f["nlogprobs"][...], dtype=np.int32)
        np.save(path[:-3] + ".npy", a)
        print(path, "->", a.shape)

if __name__ == "__main__":
    main(sys.argv[1])
