import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import spyramid_oracle as O
from semantic_pyramid_for_image_generation_b200 import models, ops
def rel(a,b): a,b=a.float().cpu(),b.float().cpu(); return float((a-b).norm()/b.norm())
ops.set_precision("split")
cf=1
v_sd=O.init_vgg_state(seed=5)
for training, mask_mode, perturb in ((True,"blob",False),(True,"inference",False),(False,"inference",False),(False,"inference",True),(False,"blob",True)):
    g_sd=O.init_generator_state(cf, seed=3)
    gen=torch.Generator().manual_seed(17)
    if perturb:
        for k,v in g_sd.items():
            if k.endswith("running_mean"): v.copy_(0.3*torch.randn(v.shape,generator=gen))
            elif k.endswith("running_var"): v.copy_(0.5+torch.rand(v.shape,generator=gen))
    images,labels,masks,z,_=O.synthetic_batch(3,seed=2,mask_mode=mask_mode)
    with torch.no_grad():
        feats=O.vgg16_features(v_sd,images)
        want=O.generator_forward({k:v.clone() for k,v in g_sd.items()},z,feats,masks,labels.float(),training=training)
    G=models.Generator(channels_factor=cf); G.load_state_dict({k:v.clone() for k,v in g_sd.items()}); G.cuda()
    G.train() if training else G.eval()
    with torch.no_grad():
        got=G(input=z.cuda(),features=[f.cuda() for f in feats],masks=[m.cuda() for m in masks],class_id=labels.float().cuda())
    print("training=%s masks=%s perturbed=%s: rel-L2 %.3e  |want| %.3e"%(training,mask_mode,perturb,rel(got,want),float(want.norm())))
