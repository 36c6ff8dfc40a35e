export ABEILLE_B200_KERNEL_TIMEOUT_S=120
mkdir -p gpurun_out
python - <<'PY'
import yaml, json, subprocess, tempfile, os, sys
ROOT=os.getcwd()
BIN=os.path.join(ROOT,"abeille_b200","lib","abl_pi_nccl")
deck=yaml.safe_load(open("tests/decks/c5g7_delta_collision_fullmesh.yaml"))
out={}
for world in (8,):
    deck["settings"].update({"nparticles": 10_000_000*world, "ngenerations": 10, "nignored": 4})
    with tempfile.NamedTemporaryFile("w",suffix=".yaml",delete=False) as f:
        yaml.safe_dump(deck,f,default_flow_style=None,sort_keys=False,width=200); path=f.name
    with tempfile.TemporaryDirectory() as td:
        idf=os.path.join(td,"id")
        ps=[subprocess.Popen(["timeout","300",BIN,path,str(r),str(world),idf,"10","4",str(r)],stdout=subprocess.PIPE,stderr=subprocess.PIPE,text=True) for r in range(world)]
        res=[p.communicate() for p in ps]
        for p,(o,e) in zip(ps,res):
            assert p.returncode==0, e[-1500:]
        j=json.loads(res[0][0].strip().splitlines()[-1])
    out[f"world_{world}"]={"nparticles":j["nparticles"],"seconds_10_generations":j["seconds"],"particles_per_s":j["nparticles"]*10/j["seconds"],"kcol_avg":j["kcol_avg"],"nbank":j["nbank"]}
print(json.dumps(out))
open("gpurun_out/t9d_nccl_pi_bench_8gpu.json","w").write(json.dumps(out))
PY
