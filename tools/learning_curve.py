"""Does a training generation learn?  B starts as a copy of A (scripts/train_iterative.py:217-219 reset_B), trains its
NoisyNet heads against the frozen A with the device rollout + PER Double-DQN updates, and its greedy win rate against A
is evaluated between rounds (eval_vs_model, :171-181).  python tools/learning_curve.py [n_envs] [rounds] [lr]"""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pingpong_selfplay_ai_b200 as pp

CFG = dict(render_size=400, paddle_width=0.2, paddle_speed=0.03, max_score=3, enable_render=False, enable_spin=True,
           magnus_factor=0.025, restitution=1, friction=0.6, ball_mass=1.0, world_ball_radius=0.03,
           ball_speed_range=[0.03, 0.05], spin_range=[-5, 5], ball_angle_intervals=[[-60, -30], [30, 60]],
           speed_scale_every=1, speed_increment=0.1)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    lr = float(sys.argv[3]) if len(sys.argv) > 3 else 2.5e-4
    prec = sys.argv[4] if len(sys.argv) > 4 else "f32"
    tui = int(sys.argv[5]) if len(sys.argv) > 5 else 200
    upc = int(sys.argv[6]) if len(sys.argv) > 6 else 16
    for seed in (0, 1, 2, 3):
        torch.manual_seed(seed); net_a = pp.QNet()
        net_b = copy.deepcopy(net_a)
        if os.environ.get("PP_LEARN_FEATURES") == "ckpt":
            # trained (frozen) feature layers of the reference's checkpoint under freshly initialised NoisyNet heads: what a
            # generation trains (scripts/train_iterative.py:97 freezes the features, :101-104 optimises the heads)
            import numpy as np
            g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "qnet_golden.npz")))
            pre = "ckpt_model5_1_fault_B/"
            sd = net_b.state_dict()
            for k in list(sd):
                if k.startswith("features"):
                    sd[k] = torch.as_tensor(g[pre + k])
            net_b.load_state_dict(sd)
        env = pp.VecPongEnv2P(n, mode="f64", serve="philox", seed=100 + seed, **CFG)
        env.reset()
        trainer = pp.DQNTrainer(net_b, batch_size=256, lr=lr, target_update_interval=tui, seed=seed)
        eng = pp.SelfPlayEngine(env, pp.Policy.qnet(net_a, noisy=True, precision=prec),
                                pp.Policy.qnet(net_b, noisy=True, eps=1.0, precision=prec), seed=seed)
        ring = pp.ReplayRing(1 << 20)
        sampler = pp.PrioritizedSampler(ring)
        wr = [pp.eval_vs_model(CFG, net_a, trainer.model, 8192, seed=5, precision=prec)]
        eps, t0 = 1.0, time.time()
        every = int(os.environ.get("PP_LEARN_EVAL_EVERY", 1))
        for r in range(rounds):
            out = pp.train_generation(eng, trainer, ring, sampler, 256, chunk=16, updates_per_chunk=upc, epsilon=eps,
                                      epsilon_decay=0.995, min_epsilon=0.02, precision=prec)
            eps = out["epsilon"]
            if (r + 1) % every == 0:
                wr.append(pp.eval_vs_model(CFG, net_a, trainer.model, 8192, seed=5, precision=prec))
        print(f"seed {seed} lr {lr} n {n} tui {tui} upc {upc}: win rate of B vs frozen A by round: " + " ".join(f"{w:.3f}" for w in wr) +
              f"  (eps {eps:.3f}, {trainer.train_steps} updates, {time.time() - t0:.1f} s)", flush=True)


if __name__ == "__main__":
    main()
