"""TEST INFRASTRUCTURE (not shipped).  Extracts the `backbone` / `classifier` blocks (and the few top-level keys a Trainer passes on) of the
reference's YAML recipes for every hot-path method into tests/golden/recipes.json, so that the GPU box — where /root/reference does not exist —
can construct every plugin exactly the way `Trainer._init_model` does (core/trainer.py:199-204, core/utils/utils.py:77-92: `get_instance`).

`benchmark` holds the SURVEY Appendix-A overrides that turn a shipped recipe into the BASELINE.json configuration (e.g. ewc.yaml ships
`resnet34` / feat_dim 512; config C1 is cifar_resnet32 / 64).  `pretrained: true` cannot be honoured offline (timm downloads the checkpoint): the test
overrides it to false (seeded random weights), the product accepts a local checkpoint through LC_B200_VIT_CHECKPOINT.

    python oracle/make_recipes.py        # needs /root/reference
"""
import glob
import json
import os

import yaml

REF = "/root/reference/config"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "recipes.json")

FILES = ["ewc.yaml", "icarl.yaml", "lwf.yaml", "finetune.yaml", "l2p-vit-cifar100-b10-10-10.yaml", "InfLoRA_opt-vit-imagenetr-b20-20-10.yaml",
         "InfLoRA_opt-vit-cifar100-b10-10-10.yaml", "InfLoRA.yaml", "dualprompt.yaml", "codaprompt.yaml", "zz_LUCIR/lucir.yaml"]
FILES += sorted(os.path.relpath(p, REF) for p in glob.glob(os.path.join(REF, "zz_SD-LoRA", "*.yaml")))
FILES += sorted(os.path.relpath(p, REF) for p in glob.glob(os.path.join(REF, "zz_GPM", "*.yaml")))

# shipped recipe -> BASELINE.json configuration
BENCH = {
    "ewc.yaml": {"config": "C1", "backbone": {"name": "cifar_resnet32"}, "classifier": {"kwargs": {"feat_dim": 64}}, "batch_size": 32},
    "icarl.yaml": {"config": "C2", "classifier": {"kwargs": {"init_cls_num": 50, "inc_cls_num": 5, "task_num": 11}}, "batch_size": 128},
    "l2p-vit-cifar100-b10-10-10.yaml": {"config": "C3", "batch_size": 128},
    "InfLoRA_opt-vit-imagenetr-b20-20-10.yaml": {"config": "C4", "batch_size": 256},
    "lwf.yaml": {"config": "C5", "backbone": {"name": "resnet18", "kwargs": {"num_classes": 200, "args": {"dataset": "tiny-imagenet", "init_cls_num": 100,
                                                                                                              "inc_cls_num": 20}}},
                 "classifier": {"kwargs": {"num_class": 200, "init_cls_num": 100, "inc_cls_num": 20}}, "batch_size": 256, "image_size": 64},
    "finetune.yaml": {"config": "-", "backbone": {"name": "cifar_resnet32"}, "classifier": {"kwargs": {"feat_dim": 64}}},
}


def main():
    out = {}
    for f in FILES:
        c = yaml.safe_load(open(os.path.join(REF, f)))
        out[f] = {"backbone": c.get("backbone"), "classifier": c.get("classifier"),
                  "top": {k: c.get(k) for k in ("init_cls_num", "inc_cls_num", "task_num", "batch_size", "image_size", "dataset", "epoch", "init_epoch")},
                  "optimizer": c.get("optimizer"), "benchmark": BENCH.get(f)}
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    print(f"wrote {OUT}: {len(out)} recipes")


if __name__ == "__main__":
    main()
