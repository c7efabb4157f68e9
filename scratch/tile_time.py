"""tiled assembly vs plain: fem2d level L with tile sizes (elements) from argv"""
import sys, os, json, subprocess
L = sys.argv[1]
for tile in sys.argv[2:]:
    env = dict(os.environ, MGB_TILE_ELEMS=tile)
    out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "te_time.py"), L, "20"], capture_output=True, text=True, env=env).stdout.strip().splitlines()[-1]
    d = json.loads(out)["times"]
    print(json.dumps(dict(L=int(L), tile_elems=int(tile), full_us=d["full"]["total_us"], full_noflush_us=d["full"]["total_noflush_us"], f0_us=d["f0"]["total_us"])), flush=True)
