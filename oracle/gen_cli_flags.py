"""Extract the reference CLI's flags (dipoorlet/__main__.py:23-55) into tests/golden/cli_flags.json by
walking the source's AST (importing the module would run parse_args() and the whole pipeline).

    python oracle/gen_cli_flags.py        # build container only; the fixture is committed
"""
import ast
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/dipoorlet/__main__.py"


def main():
    tree = ast.parse(open(SRC).read())
    flags = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
            names = [ast.literal_eval(a) for a in node.args]
            kw = {}
            for k in node.keywords:
                if k.arg == "type":
                    kw["type"] = getattr(k.value, "id", None)
                elif k.arg != "help":
                    kw[k.arg] = ast.literal_eval(k.value)
            flags.append({"names": names, **kw})
    out = os.path.join(ROOT, "tests", "golden", "cli_flags.json")
    json.dump(flags, open(out, "w"), indent=1)
    print("wrote", out, len(flags), "flags")


if __name__ == "__main__":
    main()
