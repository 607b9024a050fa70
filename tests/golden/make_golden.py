"""Regenerates the golden fixtures from the reference checkout (run in the build container only;
/root/reference does not exist on the GPU box and nothing at test time reads it).

  PpmOutputExample.txt  <- RayTracing.Test/PpmOutputExample.txt, the P3 golden file of TestPpmOutput.fs:12-46
                           (the only golden-output fixture the reference holds for this path)
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/RayTracing.Test"

if __name__ == "__main__":
    shutil.copyfile(os.path.join(REF, "PpmOutputExample.txt"), os.path.join(HERE, "PpmOutputExample.txt"))
    print("wrote PpmOutputExample.txt")
