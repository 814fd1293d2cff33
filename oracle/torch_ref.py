"""ORACLE (test infrastructure, not product): plain-PyTorch fp32 restatement of the CelebBasis hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this.
It restates, module by module and with the reference's state-dict key names, what the reference computes so
that it can travel to the GPU box (where /root/reference does not exist).  It is pinned against the
UNMODIFIED reference imported from /root/reference by oracle/make_golden.py (fixtures in tests/golden/);
tests/test_oracle_golden.py re-checks this file against those fixtures on every run.

Reference lines followed:
  UNet          ldm/modules/diffusionmodules/openaimodel.py:74-160,163-275,413-742
  attention     ldm/modules/attention.py:37-64,76-77,152-261
  util          ldm/modules/diffusionmodules/util.py:21-25,96-99,151-171,199-216
  VAE           ldm/modules/diffusionmodules/model.py:33-202,368-568 ; ldm/models/autoencoder.py:285-333
  posterior     ldm/modules/distributions/distributions.py:24-37
  CLIP text     ldm/modules/encoders/modules.py:24-31,176-298,345-406 (+ transformers CLIPEncoderLayer maths)
  embedding     ldm/modules/embedding_manager.py:279-394,452-490 ; ldm/modules/id_embedding/meta_net.py:27-87,250-346
  helpers       ldm/modules/id_embedding/helpers.py:6-41
  iresnet       ldm/modules/id_embedding/iresnet.py:26-181
  diffusion     ldm/models/diffusion/ddpm.py:126-178,289-307,590-597,926-936,1069-1116 ; ddim.py:25-204
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# =================================================================================================
# UNet
# =================================================================================================
def timestep_embedding(timesteps, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        inner = dim * mult
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim))

    def forward(self, x):
        return self.net(x)


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = dim_head * heads
        context_dim = context_dim if context_dim is not None else query_dim
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))

    def forward(self, x, context=None):
        h = self.heads
        q = self.to_q(x)
        context = context if context is not None else x
        k, v = self.to_k(context), self.to_v(context)
        b, n, _ = q.shape

        def split(t):
            return t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3).reshape(b * h, t.shape[1], -1)

        q, k, v = split(q), split(k), split(v)
        sim = torch.einsum("bid,bjd->bij", q, k) * self.scale
        attn = sim.softmax(dim=-1)
        out = torch.einsum("bij,bjd->bid", attn, v)
        out = out.reshape(b, h, n, -1).permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, context_dim=None):
        super().__init__()
        self.attn1 = CrossAttention(dim, heads=n_heads, dim_head=d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim=context_dim, heads=n_heads, dim_head=d_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, context=None):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context=context) + x
        x = self.ff(self.norm3(x)) + x
        return x


class SpatialTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, depth=1, context_dim=None):
        super().__init__()
        inner = n_heads * d_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, context_dim=context_dim) for _ in range(depth)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        x_in = x
        x = self.proj_in(self.norm(x))
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)
        for blk in self.transformer_blocks:
            x = blk(x, context=context)
        x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2)
        return self.proj_out(x) + x_in


class ResBlock(nn.Module):
    def __init__(self, channels, emb_channels, out_channels):
        super().__init__()
        self.in_layers = nn.Sequential(GroupNorm32(32, channels), nn.SiLU(), nn.Conv2d(channels, out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, out_channels))
        self.out_layers = nn.Sequential(GroupNorm32(32, out_channels), nn.SiLU(), nn.Dropout(0.0),
                                        nn.Conv2d(out_channels, out_channels, 3, padding=1))
        self.skip_connection = nn.Identity() if out_channels == channels else nn.Conv2d(channels, out_channels, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        h = h + self.emb_layers(emb).type(h.dtype)[..., None, None]
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class Downsample(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.op = nn.Conv2d(channels, channels, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class Upsample(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2, mode="nearest"))


class TimestepEmbedSequential(nn.Sequential):
    def forward(self, x, emb, context=None):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            else:
                x = layer(x)
        return x


class UNetModel(nn.Module):
    def __init__(self, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1),
                 num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768, **_):
        super().__init__()
        self.model_channels = model_channels
        ted = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, ted, mult * model_channels)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(SpatialTransformer(ch, num_heads, ch // num_heads, context_dim=context_dim))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, ted, ch), SpatialTransformer(ch, num_heads, ch // num_heads, context_dim=context_dim),
            ResBlock(ch, ted, ch))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, model_channels * mult)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(SpatialTransformer(ch, num_heads, ch // num_heads, context_dim=context_dim))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(GroupNorm32(32, ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, 3, padding=1))

    def forward(self, x, timesteps=None, context=None):
        hs = []
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels))
        h = x
        for m in self.input_blocks:
            h = m(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for m in self.output_blocks:
            h = torch.cat([h, hs.pop()], dim=1)
            h = m(h, emb, context)
        return self.out(h)


# =================================================================================================
# CLIP text transformer (maths of transformers CLIPTextModel as driven by modules.py:302-406)
# =================================================================================================
class CLIPAttention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads, self.dh = heads, dim // heads
        self.q_proj, self.k_proj = nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.v_proj, self.out_proj = nn.Linear(dim, dim), nn.Linear(dim, dim)

    def forward(self, x, mask):
        b, t, c = x.shape
        sh = lambda y: y.view(b, t, self.heads, self.dh).transpose(1, 2)
        q, k, v = sh(self.q_proj(x) * self.dh ** -0.5), sh(self.k_proj(x)), sh(self.v_proj(x))
        w = torch.matmul(q, k.transpose(-1, -2)) + mask
        w = w.softmax(dim=-1)
        o = torch.matmul(w, v).transpose(1, 2).reshape(b, t, c)
        return self.out_proj(o)


class CLIPMLP(nn.Module):
    def __init__(self, dim, inter):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(dim, inter), nn.Linear(inter, dim)

    def forward(self, x):
        x = self.fc1(x)
        return self.fc2(x * torch.sigmoid(1.702 * x))


class CLIPEncoderLayer(nn.Module):
    def __init__(self, dim, heads, inter, eps):
        super().__init__()
        self.self_attn = CLIPAttention(dim, heads)
        self.layer_norm1 = nn.LayerNorm(dim, eps=eps)
        self.mlp = CLIPMLP(dim, inter)
        self.layer_norm2 = nn.LayerNorm(dim, eps=eps)

    def forward(self, x, mask):
        x = x + self.self_attn(self.layer_norm1(x), mask)
        return x + self.mlp(self.layer_norm2(x))


class _Embeddings(nn.Module):
    def __init__(self, vocab, dim, max_pos):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, dim)
        self.position_embedding = nn.Embedding(max_pos, dim)


class _Encoder(nn.Module):
    def __init__(self, n, dim, heads, inter, eps):
        super().__init__()
        self.layers = nn.ModuleList([CLIPEncoderLayer(dim, heads, inter, eps) for _ in range(n)])


class CLIPTextTransformer(nn.Module):
    """state-dict keys match HF `CLIPTextModel().text_model` (embeddings.*, encoder.layers.N.*, final_layer_norm.*)."""

    def __init__(self, vocab=49408, dim=768, layers=12, heads=12, inter=3072, max_pos=77, eps=1e-5):
        super().__init__()
        self.embeddings = _Embeddings(vocab, dim, max_pos)
        self.encoder = _Encoder(layers, dim, heads, inter, eps)
        self.final_layer_norm = nn.LayerNorm(dim, eps=eps)

    def embed_tokens(self, ids):
        return self.embeddings.token_embedding(ids)

    def forward_embeds(self, inputs_embeds):
        """inputs_embeds: (B,T,D) token embeddings AFTER the embedding-manager rewrite, before position add."""
        b, t, _ = inputs_embeds.shape
        pos = self.embeddings.position_embedding.weight[:t]
        h = inputs_embeds + pos[None]
        mask = torch.full((t, t), torch.finfo(h.dtype).min, device=h.device, dtype=h.dtype).triu_(1)[None, None]
        for layer in self.encoder.layers:
            h = layer(h, mask)
        return self.final_layer_norm(h)


# =================================================================================================
# placeholder index arithmetic (integer path, must be bit-exact)
# =================================================================================================
def get_rep_pos(tokenized, rep_tokens):
    tok = np.asarray(tokenized)
    return [np.where(tok == t)[0] for t in rep_tokens]


def shift_index_map(d, r_pos, reps):
    """Restates helpers.py:13-41 on an index vector: returns (src, rep_final_pos_list) where src[i] is the original
    row that ends up in row i after the shift/duplicate, and the final placeholder positions per token."""
    offset = np.zeros(d, dtype=np.int64)
    r_cat = np.concatenate(r_pos) if len(r_pos) else np.zeros(0, dtype=np.int64)
    for p in r_cat:
        offset[p + 1:] += (reps - 1)
    r_cnt = r_cat.shape[0]
    target_pos = (np.arange(d) + offset)[: d - r_cnt * (reps - 1)]
    src = np.arange(d)
    src[target_pos] = np.arange(target_pos.shape[0])
    rep_final = target_pos[r_cat].repeat(reps) + np.tile(np.arange(reps), r_cnt)
    snapshot = src.copy()
    src[rep_final] = snapshot[target_pos[r_cat].repeat(reps)]
    out, lo = [], 0
    for i in range(len(r_pos)):
        times = r_pos[i].shape[0]
        out.append(rep_final[lo: lo + times * reps].reshape(times, reps))
        lo += times * reps
    return src, out


# =================================================================================================
# celeb-basis MLP + embedding rewrite
# =================================================================================================
def celeb_mlp(v, W, b, es=2):
    """meta_net.py:27-48 (lr_mul=1) + :266-273: returns normalised coefficients (F, es, 1, K)."""
    x = F.leaky_relu(F.linear(v, W, b), 0.2)
    x = x.view(v.shape[0], es, 1, -1)
    return F.normalize(x, dim=-1, p=2)


def celeb_basis(coef, basis):
    """meta_net.py:275-289: (F,es,1,K) x (es,1+K,D) -> (F, es, D)."""
    c_mean, pca = basis[:, 0], basis[:, 1:]
    z = torch.einsum("behk,ekc->behc", coef, pca) + c_mean[None, :, None, :]
    return z.reshape(coef.shape[0], -1, basis.shape[-1])


def inject_embeddings(ids, tok_emb, z, placeholder_token, reps):
    """embedding_manager.py:347-360 for num_ids == 1.  ids (B,T) int64 cpu/np, tok_emb (B,T,D), z (B, reps, D)."""
    out = []
    all_pos = []
    for b in range(tok_emb.shape[0]):
        pos = get_rep_pos(np.asarray(ids[b].cpu()), [placeholder_token])
        src, fin = shift_index_map(tok_emb.shape[1], pos, reps)
        e = tok_emb[b][torch.as_tensor(src, device=tok_emb.device)]
        rows = [e[i] for i in range(e.shape[0])]
        for one_pos in fin[0]:
            for j, p in enumerate(one_pos):
                rows[int(p)] = z[b][j]
        out.append(torch.stack(rows, 0))
        all_pos.append(fin)
    return torch.stack(out, 0), all_pos


# =================================================================================================
# diffusion schedule / loss
# =================================================================================================
def make_schedule(timesteps=1000, linear_start=0.00085, linear_end=0.0120):
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2
    betas = betas.numpy()
    alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
    return {
        "betas": torch.tensor(betas, dtype=torch.float32),
        "alphas_cumprod": torch.tensor(alphas_cumprod, dtype=torch.float32),
        "sqrt_alphas_cumprod": torch.tensor(np.sqrt(alphas_cumprod), dtype=torch.float32),
        "sqrt_one_minus_alphas_cumprod": torch.tensor(np.sqrt(1.0 - alphas_cumprod), dtype=torch.float32),
    }


def q_sample(sched, x0, t, noise):
    a = sched["sqrt_alphas_cumprod"].to(x0.device)[t].view(-1, 1, 1, 1)
    s = sched["sqrt_one_minus_alphas_cumprod"].to(x0.device)[t].view(-1, 1, 1, 1)
    return a * x0 + s * noise


def eps_loss(pred, noise):
    """ddpm.py:1084-1096 with logvar == 0, l_simple_weight 1, original_elbo_weight 0."""
    return ((pred - noise) ** 2).mean(dim=[1, 2, 3]).mean()


# =================================================================================================
# VAE encoder (AutoencoderKL.encode)
# =================================================================================================
def _swish(x):
    return x * torch.sigmoid(x)


def _vae_norm(c):
    return nn.GroupNorm(32, c, eps=1e-6, affine=True)


class VaeResnetBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1, self.conv1 = _vae_norm(cin), nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2, self.conv2 = _vae_norm(cout), nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        h = self.conv1(_swish(self.norm1(x)))
        h = self.conv2(_swish(self.norm2(h)))
        if hasattr(self, "nin_shortcut"):
            x = self.nin_shortcut(x)
        return x + h


class VaeAttnBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = _vae_norm(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))

    def forward(self, x):
        h = self.norm(x)
        q, k, v = self.q(h), self.k(h), self.v(h)
        b, c, hh, ww = q.shape
        q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
        k = k.reshape(b, c, hh * ww)
        w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
        v = v.reshape(b, c, hh * ww)
        h = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
        return x + self.proj_out(h)


class VaeDownsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class VaeEncoder(nn.Module):
    def __init__(self, *, ch, ch_mult, num_res_blocks, in_channels, z_channels, double_z=True, **_):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, padding=1)
        in_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        bin_ = ch
        for i in range(self.num_resolutions):
            lvl = nn.Module()
            lvl.block = nn.ModuleList()
            lvl.attn = nn.ModuleList()
            bin_, bout = ch * in_mult[i], ch * ch_mult[i]
            for _ in range(num_res_blocks):
                lvl.block.append(VaeResnetBlock(bin_, bout))
                bin_ = bout
            if i != self.num_resolutions - 1:
                lvl.downsample = VaeDownsample(bin_)
            self.down.append(lvl)
        self.mid = nn.Module()
        self.mid.block_1, self.mid.attn_1, self.mid.block_2 = VaeResnetBlock(bin_, bin_), VaeAttnBlock(bin_), VaeResnetBlock(bin_, bin_)
        self.norm_out = _vae_norm(bin_)
        self.conv_out = nn.Conv2d(bin_, 2 * z_channels if double_z else z_channels, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for i in range(self.num_resolutions):
            for j in range(self.num_res_blocks):
                h = self.down[i].block[j](h)
            if i != self.num_resolutions - 1:
                h = self.down[i].downsample(h)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        return self.conv_out(_swish(self.norm_out(h)))


class AutoencoderKLEncode(nn.Module):
    """encoder + quant_conv of AutoencoderKL (keys: encoder.*, quant_conv.*)."""

    def __init__(self, ddconfig, embed_dim):
        super().__init__()
        self.encoder = VaeEncoder(**ddconfig)
        self.quant_conv = nn.Conv2d(2 * ddconfig["z_channels"], 2 * embed_dim, 1)

    def forward(self, x):
        return self.quant_conv(self.encoder(x))


def posterior_sample(moments, eps, scale_factor):
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return scale_factor * (mean + torch.exp(0.5 * logvar) * eps)


# =================================================================================================
# CosFace iresnet100 + face preprocessing
# =================================================================================================
class IBasicBlock(nn.Module):
    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.bn1 = nn.BatchNorm2d(inplanes, eps=1e-5)
        self.conv1 = nn.Conv2d(inplanes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes, eps=1e-5)
        self.prelu = nn.PReLU(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes, eps=1e-5)
        self.downsample = downsample

    def forward(self, x):
        out = self.bn3(self.conv2(self.prelu(self.bn2(self.conv1(self.bn1(x))))))
        return out + (x if self.downsample is None else self.downsample(x))


class IResNet(nn.Module):
    def __init__(self, layers=(3, 13, 30, 3), num_features=512):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, eps=1e-5)
        self.prelu = nn.PReLU(64)
        self.layer1 = self._make(64, layers[0])
        self.layer2 = self._make(128, layers[1])
        self.layer3 = self._make(256, layers[2])
        self.layer4 = self._make(512, layers[3])
        self.bn2 = nn.BatchNorm2d(512, eps=1e-5)
        self.fc = nn.Linear(512 * 49, num_features)
        self.features = nn.BatchNorm1d(num_features, eps=1e-5)

    def _make(self, planes, blocks):
        ds = nn.Sequential(nn.Conv2d(self.inplanes, planes, 1, stride=2, bias=False), nn.BatchNorm2d(planes, eps=1e-5))
        layers = [IBasicBlock(self.inplanes, planes, 2, ds)]
        self.inplanes = planes
        layers += [IBasicBlock(planes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.prelu(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = torch.flatten(self.bn2(x), 1)
        return self.features(self.fc(x))


TRANS_MATRIX = [[1.07695457, -0.03625215, -1.56352194 / 512], [0.03625215, 1.07695457, -5.32134629 / 512]]


def face_preprocess(faces_nhwc):
    """meta_net.py:253-262: NHWC -> NCHW, fixed affine warp (grid_sample), bilinear resize to 112."""
    img = faces_nhwc.permute(0, 3, 1, 2)
    M = torch.tensor([TRANS_MATRIX], dtype=torch.float32, device=img.device).repeat(img.shape[0], 1, 1)
    grid = F.affine_grid(M, size=img.size(), align_corners=True)
    img = F.grid_sample(img, grid, align_corners=True, mode="bilinear", padding_mode="zeros")
    return F.interpolate(img, size=112, mode="bilinear", align_corners=True)


# =================================================================================================
# one full training-step forward (rows a1-a28 of SURVEY.md §8) on explicit weights
# =================================================================================================
class OracleModel(nn.Module):
    """Same submodule names / state-dict keys as the reference LatentDiffusion for every tensor on the path."""

    def __init__(self, params, clip_layers=12):
        super().__init__()
        self.params = params

        class _Wrap(nn.Module):
            def __init__(self, m):
                super().__init__()
                self.diffusion_model = m
        self.model = _Wrap(UNetModel(**params["unet_config"]["params"]))
        fs = params["first_stage_config"]["params"]
        self.first_stage_model = AutoencoderKLEncode(fs["ddconfig"], fs["embed_dim"])

        class _T(nn.Module):
            def __init__(self, n):
                super().__init__()
                self.text_model = CLIPTextTransformer(layers=n)

        class _C(nn.Module):
            def __init__(self, n):
                super().__init__()
                self.transformer = _T(n)
        self.cond_stage_model = _C(clip_layers)

        class _Lin(nn.Module):
            def __init__(self):
                super().__init__()
                self.weight = nn.Parameter(torch.zeros(1024, 512))
                self.bias = nn.Parameter(torch.zeros(1024))

        class _SV(nn.Module):
            def __init__(self):
                super().__init__()
                self.net = nn.Sequential(_Lin())

        class _Meta(nn.Module):
            def __init__(self):
                super().__init__()
                self.id_model = IResNet()
                self.stylegan_mlp = _SV()

        class _EM(nn.Module):
            def __init__(self):
                super().__init__()
                self.meta_id_net = _Meta()
        self.embedding_manager = _EM()
        self.sched = make_schedule(params["timesteps"], params["linear_start"], params["linear_end"])
        self.scale_factor = params["scale_factor"]

    def trainable(self):
        lin = self.embedding_manager.meta_id_net.stylegan_mlp.net[0]
        return lin.weight, lin.bias

    def step(self, batch, draws, token_ids, basis, placeholder_token):
        """Returns dict(loss, z, context, eps, x_noisy, coef, celeb_z, face_feat)."""
        out = {}
        x = batch["image"].permute(0, 3, 1, 2).contiguous().float()
        with torch.no_grad():
            moments = self.first_stage_model(x)
            z = posterior_sample(moments, draws["posterior_eps"].to(x.device), self.scale_factor)
            faces = batch["image_ori"]["faces"]
            n_id = batch["image_ori"]["ids"].shape[1]
            cat = torch.cat(faces.chunk(n_id, -1), 0)
            self.embedding_manager.meta_id_net.id_model.eval()
            v = F.normalize(self.embedding_manager.meta_id_net.id_model(face_preprocess(cat)), dim=-1, p=2)
        W, b = self.trainable()
        coef = celeb_mlp(v, W, b)
        zc = celeb_basis(coef, basis.to(x.device))
        B = x.shape[0]
        tm = self.cond_stage_model.transformer.text_model
        tok_emb = tm.embed_tokens(token_ids.to(x.device))
        emb, pos = inject_embeddings(token_ids, tok_emb, zc[:B], placeholder_token, zc.shape[1])
        context = tm.forward_embeds(emb)
        t = draws["t"].to(x.device)
        noise = draws["noise"].to(x.device)
        x_noisy = q_sample(self.sched, z, t, noise)
        eps = self.model.diffusion_model(x_noisy, t, context)
        loss = eps_loss(eps, noise)
        for tns in (emb, context, zc, coef, eps):
            if tns.requires_grad:
                tns.retain_grad()
        out.update(loss=loss, z=z, moments=moments, context=context, eps=eps, x_noisy=x_noisy, coef=coef, celeb_z=zc,
                   face_feat=v, positions=pos, emb=emb)
        return out


# =================================================================================================
# VAE decoder + DDIM (inference path: scripts/stable_txt2img.py -> ddim.py:57-204, autoencoder.py:330-333)
# =================================================================================================
class VaeUpsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class VaeDecoder(nn.Module):
    def __init__(self, *, ch, out_ch, ch_mult, num_res_blocks, z_channels, **_):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        bin_ = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, bin_, 3, padding=1)
        self.mid = nn.Module()
        self.mid.block_1, self.mid.attn_1, self.mid.block_2 = VaeResnetBlock(bin_, bin_), VaeAttnBlock(bin_), VaeResnetBlock(bin_, bin_)
        self.up = nn.ModuleList()
        for i in reversed(range(self.num_resolutions)):
            lvl = nn.Module()
            lvl.block, lvl.attn = nn.ModuleList(), nn.ModuleList()
            bout = ch * ch_mult[i]
            for _ in range(num_res_blocks + 1):
                lvl.block.append(VaeResnetBlock(bin_, bout))
                bin_ = bout
            if i != 0:
                lvl.upsample = VaeUpsample(bin_)
            self.up.insert(0, lvl)
        self.norm_out = _vae_norm(bin_)
        self.conv_out = nn.Conv2d(bin_, out_ch, 3, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        for i in reversed(range(self.num_resolutions)):
            for j in range(self.num_res_blocks + 1):
                h = self.up[i].block[j](h)
            if i != 0:
                h = self.up[i].upsample(h)
        return self.conv_out(_swish(self.norm_out(h)))


class AutoencoderKLDecode(nn.Module):
    """post_quant_conv + decoder of AutoencoderKL (keys: decoder.*, post_quant_conv.*)."""

    def __init__(self, ddconfig, embed_dim):
        super().__init__()
        self.decoder = VaeDecoder(**ddconfig)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)

    def forward(self, z):
        return self.decoder(self.post_quant_conv(z))


def ddim_sample(unet, sched, cond, uncond, x_T, steps, scale, eta=0.0):
    """ddim.py:57-204 with eta == 0 (02_start_test.sh): uniform timesteps (+1), CFG batch doubling."""
    T = sched["alphas_cumprod"].shape[0]
    c = T // steps
    ts = np.asarray(list(range(0, T, c))) + 1
    ac = sched["alphas_cumprod"].double().numpy()
    a, a_prev = ac[ts], np.asarray([ac[0]] + ac[ts[:-1]].tolist())
    x = x_T
    for i, step in enumerate(np.flip(ts)):
        idx = len(ts) - i - 1
        t = torch.full((x.shape[0],), int(step), device=x.device, dtype=torch.long)
        e_u, e_c = unet(torch.cat([x] * 2), torch.cat([t] * 2), torch.cat([uncond, cond])).chunk(2)
        e = e_u + scale * (e_c - e_u)
        at, ap = float(a[idx]), float(a_prev[idx])
        pred_x0 = (x - math.sqrt(1 - at) * e) / math.sqrt(at)
        x = math.sqrt(ap) * pred_x0 + math.sqrt(1 - ap) * e
    return x
