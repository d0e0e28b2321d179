"""Golden ids from the reference's own XTTSTokenizerFast (src/auralis/models/xttsv2/config/tokenizer.py:742-1000) on a small
synthetic BPE vocabulary.  Writes tests/golden/tokenizer_small.json (the vocabulary, a `tokenizers` file) and
tests/golden/tokenizer_ids.json (texts -> per-chunk ids).     python tests/golden/make_tokenizer_golden.py

The real vocabulary is a download; the class around it (cleaners, sentence split, [lang] prefix, [SPACE]) is what is pinned
here.  The reference was written against transformers 4.x; under 5.x `PreTrainedTokenizerFast._batch_encode_plus` is gone, so
`reference_tokenizer()` restores the one behaviour the reference relies on (no padding, no truncation: encode_batch -> ids).
Texts avoid digits for the languages whose number words come from num2words (absent here; oracle/ref_text.py stubs it)."""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

LANGS = ("en", "de", "fr", "es", "it", "pt", "pl", "zh-cn", "ar", "cs", "ru", "nl", "tr", "ja", "hu", "ko")
CASES = [
    ("en", "Hello world. This is a test of the system, Mr. Smith said; isn't it? " * 6),
    ("en", "Short one."),
    ("fr", "Bonjour le monde. Ceci est un essai, n'est-ce pas? Mme Dupont arrive. " * 6),
    ("de", "Hallo Welt! Das ist ein Versuch; z.B. heute, sagte Dr. Meier. " * 7),
    ("es", "Hola mundo. ¿Qué tal? Esto es una prueba, dijo la Sra. García. " * 6),
    ("it", "Ciao mondo. Questa è una prova. " * 10),
    ("pt", "Olá mundo. Isto é um teste, disse o Sr. Silva. " * 8),
    ("pl", "Witaj świecie. To jest test. " * 12),
    ("tr", "Merhaba dünya. Bu bir denemedir. " * 11),
    ("ru", "Привет мир. Это проверка системы. " * 11),
    ("nl", "Hallo wereld. Dit is een proef. " * 11),
    ("cs", "Ahoj světe. Toto je zkouška. " * 12),
    ("hu", "Helló világ. Ez egy próba. " * 13),
]


def build_vocab(path):
    from tokenizers import Tokenizer, models, trainers, pre_tokenizers
    specials = ["[STOP]", "[UNK]", "[SPACE]", "[START]", "[PAD]"] + [f"[{l}]" for l in LANGS]
    tok = Tokenizer(models.BPE(unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    corpus = [t.lower() for _, t in CASES]
    tok.train_from_iterator(corpus + [c.replace(" ", "[SPACE]") for c in corpus], trainers.BpeTrainer(vocab_size=400, special_tokens=specials))
    tok.save(path)


def reference_tokenizer(vocab_file):
    from oracle import ref_text
    ref = ref_text.load()
    from transformers import PreTrainedTokenizerFast as P
    if not hasattr(P, "_batch_encode_plus"):                      # transformers >= 5
        def _batch_encode_plus(self, texts, add_special_tokens=True, **kw):
            return {"input_ids": [e.ids for e in self._tokenizer.encode_batch(list(texts), add_special_tokens=add_special_tokens)]}
        P._batch_encode_plus = _batch_encode_plus
    return ref.XTTSTokenizerFast(vocab_file=vocab_file)


if __name__ == "__main__":
    vocab = os.path.join(HERE, "tokenizer_small.json")
    build_vocab(vocab)
    t = reference_tokenizer(vocab)
    out = [{"lang": lang, "text": text, "ids": [list(map(int, c)) for c in t.batch_encode_with_split(text, lang=lang)]} for lang, text in CASES]
    json.dump(out, open(os.path.join(HERE, "tokenizer_ids.json"), "w"), ensure_ascii=False)
    print(len(out), "cases;", sum(len(c["ids"]) for c in out), "chunks;", os.path.getsize(vocab), "bytes of vocabulary")
