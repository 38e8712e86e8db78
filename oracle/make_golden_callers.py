"""TEST INFRASTRUCTURE: records tests/golden/callers.json by executing the reference's own
`greedy_samples_byscore` (evaluation/inference_unigen_cot.py:507-534, extracted from the unmodified source with `ast`
so that the script's heavy module-level imports are not needed) on seeded rating tables without score ties.

    python oracle/make_golden_callers.py        # needs /root/reference and pandas
"""
import ast
import json
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("UNIGEN_REFERENCE_ROOT", "/root/reference")


def reference_fn():
    import pandas as pd
    path = os.path.join(REF, "evaluation", "inference_unigen_cot.py")
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "greedy_samples_byscore")
    ns = {"pd": pd}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["greedy_samples_byscore"]


def make_table(rng, n_inst, n_img, n_q, distinct_scores=True):
    rows = []
    for inst in rng.sample(range(1000), n_inst):
        # distinct yes-counts per image inside an instance: no ties, so the reference's order is well defined
        counts = rng.sample(range(n_q + 1), n_img) if distinct_scores else [rng.randrange(n_q + 1) for _ in range(n_img)]
        imgs = list(range(n_img))
        rng.shuffle(imgs)
        for k, yes in zip(imgs, counts):
            answers = ["yes"] * yes + ["no"] * (n_q - yes)
            rng.shuffle(answers)
            for qi, a in enumerate(answers):
                rows.append({"image_id": f"{inst:05}/samples/{k:05}.png", "prompt": f"p{inst}", "raw_answer": f"q{qi} {a}", "answer": a})
    rng.shuffle(rows)
    return rows


def main():
    fn = reference_fn()
    rng = random.Random(0)
    cases = []
    for n_inst, n_img, n_q, topk in ((1, 4, 5, 4), (3, 8, 9, 4), (5, 6, 7, 2), (2, 10, 12, 1), (4, 5, 6, 8)):
        table = make_table(rng, n_inst, n_img, n_q)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            picked, stats = fn(table, topk=topk)
        cases.append(dict(table=table, topk=topk, picked=picked,
                          stats=[dict(instance_id=r.instance_id, image_id=r.image_id, yes_count=int(r.yes_count),
                                      total_questions=int(r.total_questions), score=float(r.score)) for r in stats.itertuples()]))
    out = os.path.join(ROOT, "tests", "golden", "callers.json")
    with open(out, "w") as f:
        json.dump(cases, f)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
