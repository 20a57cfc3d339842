"""Host-side tokenizer registry the stage scripts instantiate (`str2tokenizer[args.tokenizer](args)`,
finetune/ppo.py:761-762).  The scripts only CONSTRUCT the tokenizers -- the text tower's inputs are pre-extracted
features, SURVEY.md §0 fact 2 -- so this module holds what `ppo.sh` / `pointwise.sh` select ("bpe" for the text side,
"virtual" for the ViT side) plus the two trivial ones; it is plain Python with no kernels.  "bert", "xlmroberta",
"image" and "text_image" belong to other model families and raise when selected.

BPE here is the published GPT-2 byte-level algorithm (Radford et al. 2019, `encoder.py`): bytes are mapped to
printable unicode characters, words are split with the GPT-2 regular expression and merged greedily by merge rank."""
import functools

import regex

UNK_TOKEN, CLS_TOKEN, SEP_TOKEN, MASK_TOKEN, PAD_TOKEN = "<unk>", "<s>", "</s>", "<mask>", "<pad>"


class Vocab:
    """token <-> id tables read from a one-token-per-line file (first whitespace-separated field of each line)."""

    def __init__(self):
        self.w2i, self.i2w, self.w2c = {}, [], {}

    def load(self, vocab_path, is_quiet=False):
        with open(vocab_path, mode="r", encoding="utf-8") as f:
            for i, line in enumerate(f):
                stripped = line.strip("\r\n")
                w = stripped.split()[0] if line.strip() else stripped
                self.w2i[w] = i
                self.i2w.append(w)
        if not is_quiet:
            print("Vocabulary size: ", len(self))

    def get(self, w):
        return self.w2i[w]

    def __len__(self):
        return len(self.i2w)


class Tokenizer:
    def __init__(self, args, is_src=True):
        if getattr(args, "spm_model_path" if is_src else "tgt_spm_model_path", None):
            raise ValueError("sentencepiece models are outside the LR2PPO hot path (no stage script passes one)")
        v = Vocab()
        v.load(args.vocab_path if is_src else args.tgt_vocab_path, is_quiet=True)
        self.vocab = v.w2i
        self.inv_vocab = {i: w for w, i in self.vocab.items()}

    def tokenize(self, text):
        raise NotImplementedError

    def convert_tokens_to_ids(self, tokens):
        return [self.vocab[t] for t in tokens]

    def convert_ids_to_tokens(self, ids):
        return [self.inv_vocab[i] for i in ids]


class CharTokenizer(Tokenizer):
    def tokenize(self, text, use_vocab=True):
        chars = list(text.strip())
        return [c if c in self.vocab else UNK_TOKEN for c in chars] if use_vocab else chars


class SpaceTokenizer(Tokenizer):
    def tokenize(self, text, use_vocab=True):
        words = text.strip().split(" ")
        return [w if w in self.vocab else UNK_TOKEN for w in words] if use_vocab else words


@functools.lru_cache()
def bytes_to_unicode():
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    table, extra = {b: chr(b) for b in keep}, 0
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    return table


class BPETokenizer(Tokenizer):
    SPLIT = regex.compile(r"""'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""")

    def __init__(self, args, is_src=True):
        super().__init__(args, is_src)
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {c: b for b, c in self.byte_encoder.items()}
        with open(args.merges_path if is_src else args.tgt_merges_path, encoding="utf-8") as f:
            lines = f.read().split("\n")[1:-1]                       # first line is the "#version" header
        self.bpe_ranks = {tuple(line.split()): rank for rank, line in enumerate(lines)}
        self.cache = {}

    def bpe(self, token):
        if token in self.cache:
            return self.cache[token]
        word = tuple(token)
        while len(word) > 1:
            pairs = {(a, b) for a, b in zip(word, word[1:])}
            best = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if best not in self.bpe_ranks:
                break
            merged, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and (word[i], word[i + 1]) == best:
                    merged.append(word[i] + word[i + 1])
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = tuple(merged)
        out = " ".join(word)
        self.cache[token] = out
        return out

    def tokenize(self, text):
        pieces = []
        for tok in self.SPLIT.findall(text):
            mapped = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            pieces.extend(self.bpe(mapped).split(" "))
        return pieces

    def decode(self, tokens):
        return bytearray(self.byte_decoder[c] for c in "".join(tokens)).decode("utf-8", errors="replace")


class VirtualTokenizer:
    """Placeholder for towers whose input is not text (ViT): an empty vocabulary."""

    def __init__(self, args, is_src=True):
        self.vocab = []


def _outside(name):
    def build(args, is_src=True):
        raise ValueError(f"tokenizer {name!r} belongs to another model family; the LR2PPO scripts select 'bpe' and "
                         "'virtual' (ppo.sh:43-54)")
    return build


str2tokenizer = {"char": CharTokenizer, "space": SpaceTokenizer, "bpe": BPETokenizer, "virtual": VirtualTokenizer,
                 "bert": _outside("bert"), "xlmroberta": _outside("xlmroberta"), "image": _outside("image"),
                 "text_image": _outside("text_image")}
