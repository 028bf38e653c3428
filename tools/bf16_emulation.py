"""CPU emulation of the bf16 rounding points of the device pipeline (activations, packed weights,
inter-layer gradients) on the golden training step: shows how far bf16 alone moves the gradient norms
relative to the fp32 reference (evidence for the tolerances in tests/test_gpu_model.py)."""
import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.nn.functional as F
from oracle import encoder_oracle as eo
torch.set_num_threads(8)
gold=np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'encoder_golden.npz'))
sd=eo.he_normal_state_dict(7)
class R(torch.autograd.Function):
    @staticmethod
    def forward(ctx,x): return x.to(torch.bfloat16).float()
    @staticmethod
    def backward(ctx,g): return g.to(torch.bfloat16).float()
r=R.apply
def cnn(sd,br,x):
    out=x
    for i,(name,_,_,_,_,ph,pw,pool) in enumerate(eo.CONV_SPECS):
        w=sd[f"{br}.pretrained.{name}.weight"]; b=sd[f"{br}.pretrained.{name}.bias"]
        if i>0: w=r(w)
        out=F.conv2d(out,w,b,padding=(ph,pw))
        if pool>1:
            out=r(out); out=F.max_pool2d(out,(pool,1))
        out=F.relu(out); out=r(out)
    return torch.squeeze(out,2)
def ds(sd,br,x):
    h=cnn(sd,br,x); z=F.conv1d(h,sd[f"{br}.fc1.weight"],sd[f"{br}.fc1.bias"]); e=torch.sigmoid(z); return e.reshape(e.size(0),-1)
params={k:v.clone().requires_grad_(True) for k,v in sd.items()}
batch=torch.from_numpy(gold['step_batch'])
a=ds(params,'anchor',batch[:,0:1]); p=ds(params,'postve',batch[:,1:2])
loss,cp,cn=eo.ntxent(a,p,8,0.25); loss.backward()
print('loss',float(loss),gold['train_loss_cos'])
keys=[str(k) for k in gold['layout_keys']]
for i,k in enumerate(keys):
    g=params[k].grad
    print(k, 'norm ratio %.4f'%(float(g.double().norm())/gold['grad_l2'][i]))
