#!/bin/bash
# run_nerf.py end to end on the synthetic dataset: train 3 epochs (with one quadtree refinement), checkpoint, resume,
# render_only.  Small images so that it finishes in a minute.
set -e
cd "$(dirname "$0")/../fast-learning-nerf_b200"
export FLNERF_SYN_RES=64 FLNERF_SYN_VIEWS=6
rm -rf /tmp/flnerf_logs
ARGS="--config configs/synthetic_lego.txt --basedir /tmp/flnerf_logs --expname drv --N_rand 1024 --n_epoch 4 --subdivide_every 1 --init_level 2"
python run_nerf.py $ARGS 2>&1 | grep -E "Epoch|training rays|child nodes|Saved|iter 0|Center|train complete" | head -30
ls /tmp/flnerf_logs/drv
echo "--- resume (should start at epoch 5 > n_epoch: nothing to train) and render_only"
python run_nerf.py $ARGS --n_epoch 5 2>&1 | grep -E "Reloading|load '|Epoch|training rays|Saved" | head
python run_nerf.py $ARGS --render_only --render_test 2>&1 | grep -E "RENDER|mean PSNR|Done" | head
python - <<'PY'
import torch, pickle, sys
ck = torch.load('/tmp/flnerf_logs/drv/004.tar', weights_only=False)
print(sorted(ck.keys()), len(ck['optimizer_state_dict']['state']), list(ck['network_fn_state_dict'])[:2])
sys.path.insert(0, '.')
trees = pickle.load(open('/tmp/flnerf_logs/drv/treeDivide_0004.pkl', 'rb'))
import tree
print(type(trees[0]).__module__, len(trees), [len(tree.get_children(t.root)) for t in trees], trees[0].minArea)
PY
