"""A minimal NYPC judge for the tests (INSTRUCTION.md:6-92): deals the dice, plays a random opponent that bids ANY
integer 0..100000 (not only the engine's grid), talks the stdin/stdout protocol to a JudgeSession-like `handle(line)`
callable and keeps its OWN score sheet (oracle arithmetic) to audit the bot's replies and final totals."""
import random
import re

from oracle import yacht_rules as yr

NAMES = ("ONE", "TWO", "THREE", "FOUR", "FIVE", "SIX", "CHOICE", "FOUR_OF_A_KIND", "FULL_HOUSE", "SMALL_STRAIGHT",
         "LARGE_STRAIGHT", "YACHT")
BID_RE = re.compile(r"^BID ([AB]) (\d+)$")
PUT_RE = re.compile(r"^PUT ([A-Z_]+) ([1-6]{5})$")


class Sheet:
    def __init__(self):
        self.dice, self.used, self.cats, self.bank = [], 0, [0] * 12, 0

    def total(self):
        return sum(self.cats) + (35000 if sum(self.cats[:6]) >= 63000 else 0) + self.bank

    def put(self, cat, dice):
        assert not (self.used >> cat) & 1, "category reused"
        for d in dice:
            assert d in self.dice, "die %d not held (%r)" % (d, self.dice)
            self.dice.remove(d)
        self.used |= 1 << cat
        self.cats[cat] = yr.category_points(cat, dice)


def play(handle, seed, opp_first_on_score=False, force_same_target=False):
    """Runs one full game against `handle`.  Returns (transcript, bot_sheet, opp_sheet)."""
    rng = random.Random(seed)
    bot, opp = Sheet(), Sheet()
    log = []

    def send(line, expect=None):
        reply = handle(line)
        log.append((line, reply))
        if expect is None:
            assert reply is None, (line, reply)
        return reply

    assert send("READY", True) == "OK"
    for rnd in range(1, 14):
        if rnd <= 12:
            a = [rng.randint(1, 6) for _ in range(5)]
            b = [rng.randint(1, 6) for _ in range(5)]
            m = BID_RE.match(send("ROLL %s %s" % ("".join(map(str, a)), "".join(map(str, b))), True))
            assert m, log[-1]
            g, x = m.group(1), int(m.group(2))
            assert 0 <= x <= 100000
            g0 = g if force_same_target else rng.choice("AB")
            x0 = rng.choice([0, 1, 499, 12345, 50000, 77777, 100000, rng.randint(0, 100000), x])
            if g != g0:
                bot_gets = g
            elif x != x0:
                bot_gets = g if x > x0 else ("B" if g == "A" else "A")
            else:
                bot_gets = g if rng.random() < 0.5 else ("B" if g == "A" else "A")
            opp_gets = "B" if bot_gets == "A" else "A"
            bot.dice += a if bot_gets == "A" else b
            opp.dice += a if opp_gets == "A" else b
            bot.bank += -x if bot_gets == g else x
            opp.bank += -x0 if opp_gets == g0 else x0
            send("GET %s %s %d" % (bot_gets, g0, x0))
        if rnd >= 2:
            def opp_move():
                cat = rng.choice([c for c in range(12) if not (opp.used >> c) & 1])
                dice = rng.sample(opp.dice, 5)
                opp.put(cat, list(dice))
                send("SET %s %s" % (NAMES[cat], "".join(map(str, dice))))

            def bot_move():
                m = PUT_RE.match(send("SCORE", True))
                assert m, log[-1]
                assert m.group(1) in NAMES
                bot.put(NAMES.index(m.group(1)), [int(c) for c in m.group(2)])
            for step in ((opp_move, bot_move) if opp_first_on_score else (bot_move, opp_move)):
                step()
    send("FINISH")
    assert bot.used == 4095 and opp.used == 4095 and not bot.dice and not opp.dice
    return log, bot, opp
