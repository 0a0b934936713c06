"""Stage the reference's model-definition files under baseline/_ref/ (git-ignored).

The north-star requires models/resnet_v1_5.py, efficientnet.py, deeplabv3plus.py and dcgan.py to
run UNCHANGED through the B200 backend.  /root/reference does not exist on the GPU box, so the
files are copied verbatim into the git-ignored baseline/_ref/ tree, which travels with the working
copy.  Nothing is modified and nothing is committed.
"""
import os
import shutil
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
FILES = ["models/resnet_v1_5.py", "models/resnet_v1_5_dilated.py", "models/efficientnet.py",
         "models/deeplabv3plus.py", "models/dcgan.py", "models/resnet_v1_5_wsgn.py",
         "models/efficientnet_wsgn.py"]

if not os.path.isdir(SRC):
    print("reference not found at", SRC)
    sys.exit(0)
for f in FILES:
    os.makedirs(os.path.dirname(os.path.join(DST, f)), exist_ok=True)
    shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    print("staged", f)
