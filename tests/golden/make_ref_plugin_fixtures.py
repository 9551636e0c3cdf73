"""Extracts the reference's example plugins VERBATIM into tests/golden/ref_plugins/ (test fixtures: the GPU box has no
/root/reference) and prints the normalised-source hashes that mujoco_rl_environment_wrapper_b200/plugins.py holds to
recognise them.  Run in the build container: python tests/golden/make_ref_plugin_fixtures.py

  ant_reward_function.py   benchmarking/fps_gym/fps_custom_env.py:4-27, unmodified
  readme_language.py       README.md:109-137 (the ```python block), unmodified (note: it returns a 2-tuple, which the
                           reference's own HEAD rejects, mujoco_rl.py:124,236)
  readme_reward_done.py    README.md:149-163 and 168-172, with ONE repair: the README breaks the assignment of
                           data_store[agent]["current_target"] across two lines after the `=` (a syntax error as
                           printed); the two lines are joined.  Nothing else is touched.
"""
import os
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref_plugins")
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def lines(path, lo, hi):
    return open(os.path.join(REF, path)).read().split("\n")[lo - 1:hi]


def main():
    from mujoco_rl_environment_wrapper_b200.plugins import normalised_source_hash
    os.makedirs(OUT, exist_ok=True)
    ant = "\n".join(lines("benchmarking/fps_gym/fps_custom_env.py", 4, 27)) + "\n"
    lang = "\n".join(lines("README.md", 109, 136)) + "\n"
    rd = lines("README.md", 149, 163) + [""] + lines("README.md", 168, 172)
    joined = []
    for ln in rd:
        if joined and joined[-1].rstrip().endswith('["current_target"] ='):
            joined[-1] = joined[-1].rstrip() + " " + ln.strip()
        else:
            joined.append(ln)
    rd = "\n".join(joined) + "\n"
    head = "import random\n\nimport numpy as np\n\n\n"
    for name, text in (("ant_reward_function.py", head + ant), ("readme_language.py", head + lang), ("readme_reward_done.py", head + rd)):
        open(os.path.join(OUT, name), "w").write(text)
    ns = {}
    for name in ("ant_reward_function.py", "readme_language.py", "readme_reward_done.py"):
        src = open(os.path.join(OUT, name)).read()
        exec(compile(src, name, "exec"), ns)
    import ast
    for name in ("ant_reward_function.py", "readme_language.py", "readme_reward_done.py"):
        src = open(os.path.join(OUT, name)).read()
        for node in ast.parse(src).body:
            if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
                print(node.name, normalised_source_hash(ast.get_source_segment(src, node)))


if __name__ == "__main__":
    main()
