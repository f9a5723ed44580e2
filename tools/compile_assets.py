"""Generating script for brax_tracking_b200/assets/*.npz: compiles the reference's MJCF assets with the
in-repo mini-compiler (mjcf.py), with exactly the options the reference env constructors apply
(/root/reference/envs/rodent.py:51-73, /root/reference/envs/fruitfly.py:380-413).  Needs the reference checkout."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brax_tracking_b200 import assets, mjcf  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "brax_tracking_b200", "assets")
OPT = dict(iterations=4, ls_iterations=4)  # configs/dataset/*.yaml env_args

rodent = mjcf.compile_mjcf(os.path.join(REF, "assets/rodent.xml"), scale_factor=0.9, overrides=OPT)
assets.save_model(rodent, os.path.join(OUT, "rodent.npz"))
fly = os.path.join(REF, "assets/fruitfly/fruitfly_force_fast.xml")
# six fly meshes are absent from the checkout: thorax/head keep their explicit masses with a
# sphere-equivalent inertia (declared deviation; both bodies are static in the tethered config)
ff = mjcf.compile_mjcf(fly, overrides=OPT, missing_mesh="skip")
assets.save_model(ff, os.path.join(OUT, "fly_free.npz"))
ft = mjcf.compile_mjcf(fly, delete_free_joint_of="thorax", overrides=OPT, missing_mesh="skip")
assets.save_model(ft, os.path.join(OUT, "fly_tethered.npz"))
# BASELINE.json configs[3] / SURVEY Appendix C.3 variant (i): the geometry of assets/rodent_pair.xml (two replicas of the animal,
# 114 floor contacts, 30 joint actuators per animal, no tendons) with the rodent env's rescale and solver options; the file's
# actuator block is instantiated per replica (it names un-suffixed joints and is rejected by MuJoCo as committed) and
# the stated list of inter-animal capsule-capsule pairs is opened (configs.RODENT_PAIR_ENV_ARGS)
from brax_tracking_b200 import configs  # noqa: E402
pair = mjcf.compile_mjcf(os.path.join(REF, "assets/rodent_pair.xml"), scale_factor=0.9, overrides=OPT, replicate_actuators=True,
                         extra_pairs=configs.pair_geom_names(configs.RODENT_PAIR_ENV_ARGS))
assets.save_model(pair, os.path.join(OUT, "rodent_pair.npz"))
for n, m in (("rodent", rodent), ("fly_free", ff), ("fly_tethered", ft), ("rodent_pair", pair)):
    print(n, dict(nq=m.nq, nv=m.nv, nu=m.nu, na=m.na, nbody=m.nbody, ncon=int(m.pair_ncon.sum()), nM=m.nM, cone=m.cone,
                  mass=float(m.body_mass.sum())))
