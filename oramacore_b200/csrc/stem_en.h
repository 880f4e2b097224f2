// stem_en.h — the Snowball English ("Porter2") stemming algorithm, restated from its published
// definition (snowballstem.org/algorithms/english/stemmer.html).  Host only.
//
// Why it is here: the reference tokenises AND stems every query and every indexed string through
// oramacore_lib's TextParser (token_score.rs:196-209) — an un-vendored crate (Cargo.lock:5332-5370), presumed to
// wrap the Snowball stemmers (rust-stemmers) like the rest of the Rust search ecosystem; that cannot be checked
// here, so this file pins itself to the algorithm's own published sample vocabulary (tests/test_dict_host.py)
// and is offered as the default for oc_dict_set_stemmer; a host with the reference's parser passes its own hook.
// Input: lower-case ASCII (the tokenizer's output); anything else is returned unchanged.
#pragma once
#include <cstring>
#include <string>

namespace ocs {

inline bool is_v(char c) { return c == 'a' || c == 'e' || c == 'i' || c == 'o' || c == 'u' || c == 'y'; }
inline bool ends(const std::string &w, const char *s) {
    const size_t n = strlen(s);
    return w.size() >= n && w.compare(w.size() - n, n, s) == 0;
}
// start of the region after the first non-vowel following a vowel, searching from `from`
inline size_t region_after(const std::string &w, size_t from) {
    for (size_t i = from; i + 1 < w.size(); i++)
        if (is_v(w[i]) && !is_v(w[i + 1])) return i + 2;
    return w.size();
}
inline bool has_vowel(const std::string &w, size_t end) {
    for (size_t i = 0; i < end; i++) if (is_v(w[i])) return true;
    return false;
}
// short syllable ending at position `end` (exclusive)
inline bool short_syllable_at(const std::string &w, size_t end) {
    if (end == 2) return is_v(w[0]) && !is_v(w[1]);
    if (end >= 3) {
        const char a = w[end - 3], b = w[end - 2], c = w[end - 1];
        return !is_v(a) && is_v(b) && !is_v(c) && c != 'w' && c != 'x' && c != 'Y';
    }
    return false;
}

inline std::string stem_english(const std::string &in) {
    if (in.size() <= 2) return in;
    for (char c : in) if (!((c >= 'a' && c <= 'z') || c == '\'')) return in;
    // exceptional forms
    static const char *const exc[][2] = {{"skis", "ski"}, {"skies", "sky"}, {"dying", "die"}, {"lying", "lie"}, {"tying", "tie"},
                                         {"idly", "idl"}, {"gently", "gentl"}, {"ugly", "ugli"}, {"early", "earli"}, {"only", "onli"},
                                         {"singly", "singl"}, {"sky", "sky"}, {"news", "news"}, {"howe", "howe"}, {"atlas", "atlas"},
                                         {"cosmos", "cosmos"}, {"bias", "bias"}, {"andes", "andes"}};
    for (auto &e : exc) if (in == e[0]) return e[1];
    std::string w = in;
    if (w[0] == '\'') w.erase(0, 1);
    if (w.size() <= 2) return w;
    // y -> Y where it acts as a consonant
    if (w[0] == 'y') w[0] = 'Y';
    for (size_t i = 1; i < w.size(); i++) if (w[i] == 'y' && is_v(w[i - 1])) w[i] = 'Y';
    // regions
    size_t r1;
    if (w.compare(0, 5, "gener") == 0 || w.compare(0, 5, "arsen") == 0) r1 = 5;
    else if (w.compare(0, 6, "commun") == 0) r1 = 6;
    else r1 = region_after(w, 0);
    size_t r2 = region_after(w, r1);
    auto in_r1 = [&](size_t suffix_len) { return w.size() - suffix_len >= r1; };
    auto in_r2 = [&](size_t suffix_len) { return w.size() - suffix_len >= r2; };
    auto cut = [&](size_t n) { w.erase(w.size() - n); };
    auto repl = [&](size_t n, const char *by) { w.erase(w.size() - n); w += by; };
    // step 0
    if (ends(w, "'s'")) cut(3); else if (ends(w, "'s")) cut(2); else if (ends(w, "'")) cut(1);
    // step 1a
    if (ends(w, "sses")) repl(4, "ss");
    else if (ends(w, "ied") || ends(w, "ies")) repl(3, w.size() > 4 ? "i" : "ie");
    else if (ends(w, "us") || ends(w, "ss")) {}
    else if (ends(w, "s")) { if (w.size() >= 2 && has_vowel(w, w.size() - 2)) cut(1); }
    static const char *const inv[] = {"inning", "outing", "canning", "herring", "earring", "proceed", "exceed", "succeed"};
    for (auto *e : inv) if (w == e) return w;
    // step 1b
    {
        bool did = false;
        if (ends(w, "eedly")) { if (in_r1(5)) repl(5, "ee"); }
        else if (ends(w, "eed")) { if (in_r1(3)) repl(3, "ee"); }
        else {
            size_t n = 0;
            if (ends(w, "ingly")) n = 5; else if (ends(w, "edly")) n = 4; else if (ends(w, "ing")) n = 3; else if (ends(w, "ed")) n = 2;
            if (n && has_vowel(w, w.size() - n)) { cut(n); did = true; }
        }
        if (did) {
            if (ends(w, "at") || ends(w, "bl") || ends(w, "iz")) w += 'e';
            else if (w.size() >= 2 && w[w.size() - 1] == w[w.size() - 2] && strchr("bdfgmnprt", w.back())) cut(1);
            else if (r1 >= w.size() && short_syllable_at(w, w.size())) w += 'e';   // short word
        }
    }
    // step 1c
    if (w.size() > 2 && (w.back() == 'y' || w.back() == 'Y') && !is_v(w[w.size() - 2])) w.back() = 'i';
    // step 2 (longest suffix first), in R1
    {
        static const char *const s2[][2] = {{"ization", "ize"}, {"ational", "ate"}, {"fulness", "ful"}, {"ousness", "ous"}, {"iveness", "ive"},
                                            {"tional", "tion"}, {"biliti", "ble"}, {"lessli", "less"}, {"entli", "ent"}, {"ation", "ate"},
                                            {"alism", "al"}, {"aliti", "al"}, {"ousli", "ous"}, {"iviti", "ive"}, {"fulli", "ful"},
                                            {"enci", "ence"}, {"anci", "ance"}, {"abli", "able"}, {"izer", "ize"}, {"ator", "ate"},
                                            {"alli", "al"}, {"bli", "ble"}};
        bool done = false;
        for (auto &e : s2)
            if (ends(w, e[0])) { if (in_r1(strlen(e[0]))) repl(strlen(e[0]), e[1]); done = true; break; }
        if (!done) {
            if (ends(w, "ogi")) { if (in_r1(3) && w.size() >= 4 && w[w.size() - 4] == 'l') cut(1); }
            else if (ends(w, "li")) { if (in_r1(2) && w.size() >= 3 && strchr("cdeghkmnrt", w[w.size() - 3])) cut(2); }
        }
    }
    // step 3, in R1
    {
        static const char *const s3[][2] = {{"ational", "ate"}, {"tional", "tion"}, {"alize", "al"}, {"icate", "ic"}, {"iciti", "ic"},
                                            {"ical", "ic"}, {"ness", ""}, {"ful", ""}};
        bool done = false;
        for (auto &e : s3)
            if (ends(w, e[0])) { if (in_r1(strlen(e[0]))) repl(strlen(e[0]), e[1]); done = true; break; }
        if (!done && ends(w, "ative") && in_r1(5) && in_r2(5)) cut(5);
    }
    // step 4, in R2
    {
        static const char *const s4[] = {"ement", "ance", "ence", "able", "ible", "ment", "ant", "ent", "ism", "ate", "iti", "ous", "ive", "ize",
                                         "al", "er", "ic"};
        bool done = false;
        for (auto *e : s4)
            if (ends(w, e)) { if (in_r2(strlen(e))) cut(strlen(e)); done = true; break; }
        if (!done && ends(w, "ion") && in_r2(3) && w.size() >= 4 && (w[w.size() - 4] == 's' || w[w.size() - 4] == 't')) cut(3);
    }
    // step 5
    if (ends(w, "e")) { if (in_r2(1) || (in_r1(1) && !short_syllable_at(w, w.size() - 1))) cut(1); }
    else if (ends(w, "l")) { if (in_r2(1) && w.size() >= 2 && w[w.size() - 2] == 'l') cut(1); }
    for (char &c : w) if (c == 'Y') c = 'y';
    return w;
}

}  // namespace ocs
