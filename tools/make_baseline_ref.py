"""Stage the UNMODIFIED reference checkout as baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box
with the snapshot, where /root/reference does not exist).  Used by the `-m gpu` drop-in tests (the reference's own
inference.py / train.py driven through michigan_b200.launch) and by `bench.py --impl reference` (the reference's CPU
path).  Nothing from it is committed; the product never imports it.

    python tools/make_baseline_ref.py [--src /root/reference]
"""
import argparse
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default=os.environ.get("MICHIGAN_REFERENCE", "/root/reference"))
    a = ap.parse_args()
    dst = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(a.src, "models", "networks")):
        raise SystemExit("reference not found at %s" % a.src)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copytree(a.src, dst, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc", "ui", "ui_util", "demo.py", "teaser.jpg"))
    # a 3-sample training set in the layout data/custom_dataset.py expects, from the reference's own demo files
    demo = os.path.join(dst, "datasets", "FFHQ_demo")
    tr = os.path.join(dst, "datasets", "FFHQ_demo_train")
    for sub, srcsub in (("train_labels", "labels"), ("train_images", "images"), ("train_dense_orients", "orients")):
        os.makedirs(os.path.join(tr, sub), exist_ok=True)
        for f in sorted(os.listdir(os.path.join(demo, srcsub))):
            shutil.copy(os.path.join(demo, srcsub, f), os.path.join(tr, sub, f))
    n = sum(len(fs) for _, _, fs in os.walk(dst))
    print("staged %d files under %s" % (n, dst))


if __name__ == "__main__":
    main()
